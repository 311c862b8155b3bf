"""cycles per tcgen05.mma (bring-up probe): variants 0 SS N=256, 1 SS N=128, 2 SS N=64, 3 TS N=64, 4 TS N=128,
5 TS N=256, 6 TS N=64 alternating accumulators; 1 CTA alone and 148 CTAs concurrently."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnt_speech_recognition_b200 import _lib

L = _lib.load()
names = ["SS N=256", "SS N=128", "SS N=64", "TS N=64", "TS N=128", "TS N=256", "TS N=64 alt-acc", "TS N=64 walking A+B (20/block)", "TS N=64 walking B only", "walking + commit/40", "walking + commit/40 + LDTM warps"]
floor = [128, 64, 32, 32, 64, 128, 32, 32, 32, 32, 32]
out = torch.zeros(148, device="cuda")
for ctas in (1, 148):
    for v, n in enumerate(names):
        rc = L.rnntb200_debug_mma_probe(v, 4000, ctas, C.c_void_p(out.data_ptr()))
        o = out[:ctas].cpu()
        print("ctas %3d  %-16s rc %d  cycles/MMA mean %.1f max %.1f  (floor %d)" % (ctas, n, rc, o.mean().item(), o.max().item(), floor[v]))
