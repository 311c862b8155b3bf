import os, sys, torch
sys.path.insert(0, '/root/repo')
import rnnt_speech_recognition_b200 as rb
from rnnt_speech_recognition_b200 import _lib
B,T,U,V,H = 32,512,128,1024,640
g = torch.Generator().manual_seed(1234)
t = [x.cuda() for x in (torch.randn(B,T,H,generator=g), torch.randn(B,U,H,generator=g), torch.randn(H,V,generator=g)/H**0.5, torch.zeros(V))]
lab = torch.randint(1,V,(B,U-1),generator=g,dtype=torch.int32).cuda()
il, ll = torch.full((B,),T,dtype=torch.int32).cuda(), torch.full((B,),U-1,dtype=torch.int32).cuda()
for keep in (False, True):
    for _ in range(3):
        c = rb.joint_rnnt_loss(*t, lab, il, ll, precision="bf16", keep_activations=keep)
    torch.cuda.synchronize()
    _lib.set_timing(True)
    for _ in range(5):
        c = rb.joint_rnnt_loss(*t, lab, il, ll, precision="bf16", keep_activations=keep)
    torch.cuda.synchronize()
    d = {}
    for n, ms in _lib.get_timings(): d.setdefault(n, []).append(ms)
    _lib.set_timing(False)
    print("FWD=%s DBG=%s keep=%d" % (os.environ.get("RNNTB200_FWD"), os.environ.get("RNNTB200_DBG"), keep), {k: round(sum(v)/len(v),3) for k,v in d.items() if "joint" in k})
