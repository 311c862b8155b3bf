"""Bring-up diagnostics for the tcgen05 path: run fp32-exact and bf16-TC on the same input and print where
they diverge (lse plane, cached log-prob planes, costs, gradients).  Usage: python tools/debug_tc.py [B T U V H]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnnt_speech_recognition_b200 import _lib  # noqa: E402


def planes(ws, B, T, U):
    f = ws.view(torch.float32)
    SK, N = (T + U - 1) * U, T * U
    o = 0
    out = {}
    for n, sz in (("lpb", B * SK), ("lpl", B * SK), ("alphas", B * SK), ("betas", B * SK), ("lse", B * N),
                  ("llf", B), ("llb", B)):
        out[n] = f[o:o + sz].clone()
        o += sz
    return out


def run(prec, enc, pred, W, b, lab, il, ll):
    L = _lib.load()
    B, T, H = enc.shape
    U, V = pred.shape[1], W.shape[1]
    d = _lib.JointDesc(B, T, U, H, V, 0, prec, torch.cuda.current_stream().cuda_stream)
    sz = C.c_size_t(0)
    assert L.rnntb200_joint_workspace_size(C.byref(d), C.byref(sz)) == 0
    ws = torch.zeros(sz.value, dtype=torch.uint8, device="cuda")
    costs = torch.zeros(B, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    st = L.rnntb200_joint_loss_forward(C.byref(d), p(enc), p(pred), p(W), p(b), p(lab), p(ll), p(il), p(costs), p(ws))
    torch.cuda.synchronize()
    print("forward status", st, "prec", prec)
    pl = planes(ws, B, T, U)
    gc = torch.full((B,), 1.0 / B, device="cuda")
    outs = [torch.zeros_like(t) for t in (enc, pred, W, b)]
    st = L.rnntb200_joint_loss_backward(C.byref(d), p(enc), p(pred), p(W), p(b), p(lab), p(ll), p(il), p(gc),
                                        *(p(o) for o in outs), p(ws))
    torch.cuda.synchronize()
    print("backward status", st)
    return costs, pl, outs


def main():
    B, T, U, V, H = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 6 else (2, 20, 8, 64, 64)
    torch.manual_seed(0)
    enc, pred = torch.randn(B, T, H, device="cuda"), torch.randn(B, U, H, device="cuda")
    W, b = torch.randn(H, V, device="cuda") / H ** 0.5, 0.1 * torch.randn(V, device="cuda")
    lab = torch.randint(1, V, (B, max(U - 1, 1)), dtype=torch.int32, device="cuda")
    il = torch.full((B,), T, dtype=torch.int32, device="cuda")
    ll = torch.full((B,), U - 1, dtype=torch.int32, device="cuda")
    c0, p0, g0 = run(0, enc, pred, W, b, lab, il, ll)
    c1, p1, g1 = run(1, enc, pred, W, b, lab, il, ll)
    print("costs fp32", c0.tolist()[:4], "\ncosts bf16", c1.tolist()[:4])
    for n in ("lse", "llf", "llb"):
        d = (p0[n] - p1[n]).abs()
        print("%-6s max|diff| %.4g  mean %.4g  (ref max %.4g)  nan %d" % (n, d.max().item(), d.mean().item(),
              p0[n].abs().max().item(), torch.isnan(p1[n]).sum().item()))
    lse0, lse1 = p0["lse"].view(B, T, U), p1["lse"].view(B, T, U)
    print("lse signed mean (bf16-fp32) %.4g std %.4g" % ((lse1 - lse0).mean().item(), (lse1 - lse0).std().item()))
    print("lse fp32 [0,0,:8]", lse0[0, 0, :8].tolist())
    print("lse bf16 [0,0,:8]", lse1[0, 0, :8].tolist())
    SK = (T + U - 1) * U
    for n in ("lpb", "lpl"):
        a0, a1 = p0[n].view(B, T + U - 1, U), p1[n].view(B, T + U - 1, U)
        mask = torch.zeros_like(a0, dtype=torch.bool)
        for t in range(T):
            for u in range(U if n == "lpb" else U - 1):
                mask[:, t + u, u] = True
        d = (a0 - a1).abs()[mask]
        sd = (a1 - a0)[mask]
        print("%-6s max|diff| %.4g mean|diff| %.4g  signed mean (bf16-fp32) %.4g  std %.4g" % (n, d.max().item(), d.mean().item(), sd.mean().item(), sd.std().item()))
    for n, a, bb in zip(("d_enc", "d_pred", "dW", "db"), g0, g1):
        d = (a - bb).abs()
        print("%-6s max|diff| %.4g  ref max %.4g  rel-fro %.4g nan %d" % (n, d.max().item(), a.abs().max().item(),
              ((a - bb).norm() / a.norm()).item(), torch.isnan(bb).sum().item()))


if __name__ == "__main__":
    main()
