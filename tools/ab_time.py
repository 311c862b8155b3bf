"""alpha/beta wavefront kernel time at the C3 lattice (and the whole forward call)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnnt_speech_recognition_b200 as rb
from rnnt_speech_recognition_b200 import _lib
B, T, U, V, H = 32, 512, 128, 1024, 640
g = torch.Generator().manual_seed(1)
t = [x.cuda() for x in (torch.randn(B, T, H, generator=g), torch.randn(B, U, H, generator=g), torch.randn(H, V, generator=g) / 25, torch.zeros(V))]
lab = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).cuda()
il, ll = torch.full((B,), T, dtype=torch.int32).cuda(), torch.full((B,), U - 1, dtype=torch.int32).cuda()
for _ in range(3):
    rb.joint_rnnt_loss(*t, lab, il, ll, precision="bf16")
torch.cuda.synchronize(); _lib.set_timing(True)
for _ in range(5):
    rb.joint_rnnt_loss(*t, lab, il, ll, precision="bf16")
torch.cuda.synchronize()
d = {}
for n, ms in _lib.get_timings():
    d.setdefault(n, []).append(ms)
print({k: round(sum(v) / len(v), 4) for k, v in d.items()})
