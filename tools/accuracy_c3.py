"""Measured error of the bf16 tensor-core path against the fp32 exact CUDA path at the benchmark configuration
(BASELINE C3 by default): costs and all four gradients.  `python tools/accuracy_c3.py [B T U V H]`"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnnt_speech_recognition_b200 as rb

B, T, U, V, H = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 6 else (32, 512, 128, 1024, 640)
g = torch.Generator().manual_seed(1234)
enc, pred = torch.randn(B, T, H, generator=g).cuda(), torch.randn(B, U, H, generator=g).cuda()
W, b = (torch.randn(H, V, generator=g) / H ** 0.5).cuda(), torch.zeros(V).cuda()
lab = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).cuda()
il = torch.full((B,), T, dtype=torch.int32).cuda()
ll = torch.full((B,), U - 1, dtype=torch.int32).cuda()


def run(prec):
    t = [x.clone().requires_grad_() for x in (enc, pred, W, b)]
    costs = rb.joint_rnnt_loss(*t, lab, il, ll, precision=prec)
    (costs.sum() / B).backward()
    torch.cuda.synchronize()
    return costs.detach().double(), [x.grad.double() for x in t]


c32, g32 = run("fp32")
c16, g16 = run("bf16")
out = {"config": [B, T, U, V, H], "cost_mean": c32.mean().item(),
       "cost_max_rel_err": ((c16 - c32).abs() / c32.abs()).max().item()}
for n, a, r in zip(("d_enc", "d_pred", "dW", "db"), g16, g32):
    out[n] = {"rel_frobenius": ((a - r).norm() / r.norm()).item(), "max_abs_err": (a - r).abs().max().item(),
              "max_abs_ref": r.abs().max().item()}
print(json.dumps(out))
