"""Locate non-finite rows in the bf16 backward workspace (dl / zb / dz) for a ragged batch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from rnnt_speech_recognition_b200 import _lib  # noqa: E402
from test_gpu_joint import synth  # noqa: E402


def al(x, a=256):
    return (x + a - 1) // a * a


def main():
    B, T, U, V, H, seed = [int(x) for x in sys.argv[1:7]] if len(sys.argv) >= 7 else (3, 37, 19, 128, 128, 1)
    k = synth(B, T, U, V, H, seed, True)
    dev = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).cuda()
    enc, pred, W, b = (dev(k[n], torch.float32) for n in ("enc", "pred", "W", "b"))
    lab, il, ll = (dev(k[n], torch.int32) for n in ("labels", "input_lengths", "label_lengths"))
    print("lengths T", k["input_lengths"], "U", k["label_lengths"] + 1)
    L = _lib.load()
    d = _lib.JointDesc(B, T, U, H, V, 0, 1, torch.cuda.current_stream().cuda_stream)
    sz = C.c_size_t(0)
    L.rnntb200_joint_workspace_size(C.byref(d), C.byref(sz))
    ws = torch.full((sz.value,), 0xFF, dtype=torch.uint8, device="cuda")   # NaN pattern everywhere
    costs = torch.zeros(B, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    print("fwd", L.rnntb200_joint_loss_forward(C.byref(d), p(enc), p(pred), p(W), p(b), p(lab), p(ll), p(il), p(costs), p(ws)))
    gc = torch.full((B,), 1.0 / B, device="cuda")
    outs = [torch.zeros_like(t) for t in (enc, pred, W, b)]
    print("bwd", L.rnntb200_joint_loss_backward(C.byref(d), p(enc), p(pred), p(W), p(b), p(lab), p(ll), p(il), p(gc),
                                                *(p(o) for o in outs), p(ws)))
    torch.cuda.synchronize()
    for n, o in zip(("d_enc", "d_pred", "dW", "db"), outs):
        print(n, "nan:", torch.isnan(o).sum().item(), "of", o.numel())
    # workspace layout (rnnt_b200.cu JointWs, joint_tc.cuh tc_scratch_layout)
    SK, N = (T + U - 1) * U, B * T * U
    off = al(B * (4 * SK + T * U + 2) * 4) + al(N * 16)
    uu, best = 128, 1 << 30
    UU = 128
    while uu >= 8:
        pad = (U + uu - 1) // uu * uu
        if pad < best:
            best, UU = pad, uu
        uu >>= 1
    TT = 128 // UU
    nTb, nUb = (T + TT - 1) // TT, (U + UU - 1) // UU
    rows = B * nTb * nUb * 128
    print("tile %dx%d nTb %d nUb %d rows %d" % (TT, UU, nTb, nUb, rows))
    o = off + 2 * al(V * H * 2)
    dl = ws[o:o + rows * V * 2].view(torch.bfloat16).view(rows, V)
    o += al(rows * V * 2)
    zb = ws[o:o + rows * H * 2].view(torch.bfloat16).view(rows, H)
    for name, a in (("dl", dl), ("zb", zb)):
        bad = (~torch.isfinite(a.float())).any(dim=1)
        idx = bad.nonzero().flatten()
        print(name, "rows with non-finite:", idx.numel(), "tiles:", sorted(set((idx // 128).tolist()))[:40])
        if idx.numel():
            r = idx[0].item()
            print("  first bad row", r, "tile", r // 128, "r", r % 128, "cols bad", (~torch.isfinite(a[r].float())).nonzero().flatten()[:10].tolist())
    per = nTb * nUb
    inval = []
    for t in range(B * per):
        bb, rem = divmod(t, per)
        t0, u0 = (rem // nUb) * TT, (rem % nUb) * UU
        if not (t0 < k["input_lengths"][bb] and u0 < k["label_lengths"][bb] + 1):
            inval.append(t)
    print("invalid tiles:", inval[:60])


if __name__ == "__main__":
    main()
