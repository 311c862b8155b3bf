"""Bring-up profile of the two backward kernels (RNNTB200_PROF=1): cycles each warp role spent waiting on each of its
barriers, averaged over the CTAs, for one C3 forward+backward.  `python tools/role_profile.py`"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["RNNTB200_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnnt_speech_recognition_b200 as rb
from rnnt_speech_recognition_b200 import _lib

B, T, U, V, H = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 6 else (32, 512, 128, 1024, 640)
g = torch.Generator().manual_seed(1234)
t = [x.cuda().requires_grad_() for x in (torch.randn(B, T, H, generator=g), torch.randn(B, U, H, generator=g),
                                         torch.randn(H, V, generator=g) / H ** 0.5, torch.zeros(V))]
lab = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).cuda()
il, ll = torch.full((B,), T, dtype=torch.int32).cuda(), torch.full((B,), U - 1, dtype=torch.int32).cuda()
for _ in range(2):
    costs = rb.joint_rnnt_loss(*t, lab, il, ll, precision="bf16")
    (costs.sum() / B).backward()
torch.cuda.synchronize()
L = _lib.load()
L.rnntb200_debug_prof.restype = C.POINTER(C.c_longlong)
names = {0: ("bwd_dz_kernel", ["TMA: stage_empty", "MMA: priv_free / stage_full / shared_free", "epilogue warp: acc_full / named barrier"]),
         1: ("bwd_dw2_kernel (bwd_dw_kernel with RNNTB200_DW=1)", ["TMA: stage_empty", "MMA: b_full / a_ready", "producer warp: stage_empty / acc_full / rs_full"])}
# CTA-pair kernels: only the leader's MMA warp works, so the MMA row (an average over all CTAs) shows HALF the leader's numbers
for which in (0, 1):
    ptr = L.rnntb200_debug_prof(which)
    a = np.ctypeslib.as_array(ptr, shape=(256, 4, 8)).copy()
    a = a[a[:, 1, 0] > 0]
    print(names[which][0], "CTAs", len(a))
    for role in range(3):
        m = a[:, role, :5].mean(axis=0)
        print("  %-60s total %10.0f  waits %s" % (names[which][1][role], m[0], " ".join("%10.0f" % x for x in m[1:])))
