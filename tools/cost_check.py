"""Three routes to the costs at a large shape: fp32 exact fused path, torch fp32 logits + op path, bf16 fused path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnnt_speech_recognition_b200 as rb

B, T, U, V, H = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 6 else (8, 512, 128, 1024, 640)
g = torch.Generator().manual_seed(1234)
enc, pred = torch.randn(B, T, H, generator=g).cuda(), torch.randn(B, U, H, generator=g).cuda()
W, b = (torch.randn(H, V, generator=g) / H ** 0.5).cuda(), torch.zeros(V).cuda()
lab = torch.randint(1, V, (B, U - 1), generator=g, dtype=torch.int32).cuda()
il = torch.full((B,), T, dtype=torch.int32).cuda()
ll = torch.full((B,), U - 1, dtype=torch.int32).cuda()
with torch.no_grad():
    a = rb.joint_rnnt_loss(enc, pred, W, b, lab, il, ll, precision="fp32")
    logits = torch.tanh(enc[:, :, None] + pred[:, None]) @ W + b
    bb = rb.rnnt_loss(logits, lab, il, ll)
    l64 = (torch.tanh(enc.double()[:, :, None] + pred.double()[:, None]) @ W.double() + b.double())
    c = rb.joint_rnnt_loss(enc, pred, W, b, lab, il, ll, precision="bf16")
    # fp64 reference through the fp64 C-ABI entry would need 2x memory; compare the two fp32 routes instead
torch.cuda.synchronize()
print("fp32 fused :", [round(x, 3) for x in a.tolist()])
print("fp32 op    :", [round(x, 3) for x in bb.tolist()])
print("bf16 fused :", [round(x, 3) for x in c.tolist()])
print("max |fused32 - op32| =", (a - bb).abs().max().item(), " max |bf16 - op32| =", (c - bb).abs().max().item())
print("logits fp32 vs fp64 max abs:", (logits.double() - l64).abs().max().item())
