"""Bring-up probe of the 2-CTA MMA form (tools/probes/mma2_probe.cuh, built by tools/probes/build_probe.sh into tools/probes/libprobe.so): `timeout 60 python tools/probes/mma2_probe.py`.

mode 0: every CTA of every pair dumps its 128 x 64 accumulator of D = A_cta . B^T with small-integer operands generated
        in the kernel; compared EXACTLY with the integer GEMM below.  A mismatch pattern tells which assumption of the
        round-2 plan is wrong (B split across the pair by vocabulary rows [32*rank, +32); same TMEM columns in both CTAs).
mode 1: cycles per cta_group::2 MMA with 20 MMAs per elected block, 1 pair alone and 74 pairs concurrently
        (floor: 32 cycles for M=256 N=64 K=16, the same as the 1-CTA M=128 form per SM).
Run it under `timeout`: a wrong barrier protocol shows up as a hang."""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libprobe.so")
if not os.path.exists(SO):
    import subprocess
    subprocess.run([os.path.join(HERE, "build_probe.sh")], check=True)
L = C.CDLL(SO)
KB, N = 5, 64
K = KB * 64


def probe_a(cta, r, k):
    return ((r * 3 + k * 5 + cta * 7) % 7) - 3


def probe_b(n, k):
    return ((n * 2 + k) % 5) - 2


r, k, n = np.arange(128)[:, None], np.arange(K)[None, :], np.arange(N)[:, None]
Bm = probe_b(n, k).astype(np.float64)                                   # (N, K)
want = np.stack([probe_a(c, r, k).astype(np.float64) @ Bm.T for c in (0, 1)])   # (2, 128, N)

for clusters in (1, 74):
    out = torch.full((clusters, 2, 128, N), float("nan"), device="cuda")
    rc = L.rnntb200_wip_mma2_probe(0, 0, clusters, C.c_void_p(out.data_ptr()))
    got = out.cpu().numpy().astype(np.float64)
    bad = int((got != want[None]).sum())
    print("mode 0  clusters %3d  rc %d  mismatching accumulator entries: %d of %d" % (clusters, rc, bad, got.size))
    if bad:
        c, cta, row, col = np.argwhere(got != want[None])[0]
        print("   first mismatch at cluster %d cta %d row %d col %d: got %r want %r" % (c, cta, row, col, got[c, cta, row, col], want[cta, row, col]))
        # a common wrong assumption, for the record
        alt = np.stack([probe_a(c_, r, k).astype(np.float64) @ np.concatenate([Bm[32 * c_:32 * c_ + 32]] * 2).T for c_ in (0, 1)])
        print("   matches 'each CTA only sees its own B half': %s" % bool((got[0] == alt).all()))

t = torch.zeros(74, device="cuda")
for clusters in (1, 74):
    rc = L.rnntb200_wip_mma2_probe(1, 4000, clusters, C.c_void_p(t.data_ptr()))
    o = t[:clusters].cpu()
    print("mode 1  clusters %3d  rc %d  cycles per 2-CTA MMA mean %.1f max %.1f (floor 32)" % (clusters, rc, o.mean().item(), o.max().item()))
