"""CUDA-graph capture of one fused forward+backward (sync-free configuration, compact=False) and replay check."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnnt_speech_recognition_b200 as rb

torch.manual_seed(0)
B, T, U, V, H = 4, 64, 32, 256, 128
enc = torch.randn(B, T, H, device="cuda", requires_grad=True)
pred = torch.randn(B, U, H, device="cuda", requires_grad=True)
W = (torch.randn(H, V, device="cuda") / H ** 0.5).requires_grad_()
b = torch.zeros(V, device="cuda", requires_grad=True)
lab = torch.randint(1, V, (B, U - 1), dtype=torch.int32, device="cuda")
il = torch.tensor([T, T - 7, 20, T], dtype=torch.int32, device="cuda")
ll = torch.tensor([U - 1, 5, U - 1, 12], dtype=torch.int32, device="cuda")


def step():
    for t in (enc, pred, W, b):
        t.grad = None
    costs = rb.joint_rnnt_loss(enc, pred, W, b, lab, il, ll, precision="bf16", compact=False)
    (costs.sum() / B).backward()
    return costs


for _ in range(3):   # warm-up: builds the library handles / side stream outside the capture
    ref = step()
torch.cuda.synchronize()
ref_costs, ref_g = ref.detach().clone(), [t.grad.clone() for t in (enc, pred, W, b)]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        out = step()
torch.cuda.current_stream().wait_stream(s)
with torch.no_grad():
    enc.mul_(1.0)   # same inputs: replay must reproduce the eager result
g.replay()
torch.cuda.synchronize()
print("graph replay costs equal:", torch.equal(out, ref_costs))
for n, t, r in zip(("d_enc", "d_pred", "dW", "db"), (enc, pred, W, b), ref_g):
    print(n, "max|diff| vs eager", (t.grad - r).abs().max().item())
