"""N identical forward+backward steps of the tensor-core path at a BASELINE shape (default C3), nothing else: the target of
the ncu captures under profiles/ (`ncu -k regex:rb:: ... python tools/one_step.py 3`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnnt_speech_recognition_b200 as rb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, T, U, V, H = [int(x) for x in sys.argv[2:7]] if len(sys.argv) >= 7 else (32, 512, 128, 1024, 640)
g = torch.Generator().manual_seed(1234)
t = [x.cuda().requires_grad_() for x in (torch.randn(B, T, H, generator=g), torch.randn(B, U, H, generator=g),
                                         torch.randn(H, V, generator=g) / H ** 0.5, torch.zeros(V))]
lab = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).cuda()
il, ll = torch.full((B,), T, dtype=torch.int32).cuda(), torch.full((B,), U - 1, dtype=torch.int32).cuda()
for _ in range(steps):
    for x in t:
        x.grad = None
    costs = rb.joint_rnnt_loss(*t, lab, il, ll, precision="bf16")
    (costs.sum() / B).backward()
torch.cuda.synchronize()
print("ok", costs[:2].tolist())
