"""Bring-up check of the tensor-core path (csrc/joint_tc4.cuh, bwd_tc.cuh): runs a list of shapes through joint_rnnt_loss in
keep and in recompute mode and compares costs and every gradient with the fp32 exact CUDA path of the same library, and the
two modes with each other (they must agree bit for bit).  Quicker than the pytest suite while a kernel is being changed;
combine with RNNTB200_FWD / RNNTB200_DZ / RNNTB200_DW to check the one-CTA forms.

    timeout 600 python tools/bwd_check.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [  # B, T, U, V, H, seed, ragged
    (2, 20, 8, 64, 64, 0, False),
    (2, 20, 8, 256, 128, 7, False),
    (3, 37, 19, 128, 128, 1, True),
    (2, 50, 40, 192, 320, 2, True),
    (1, 9, 140, 64, 64, 3, False),
    (2, 33, 128, 512, 640, 4, True),
    (4, 100, 50, 1024, 640, 5, True),
]
NAMES = ("d_enc", "d_pred", "dW", "db")


def run_all(keep_modes=(True, False)):
    from test_gpu_joint import run_joint, synth
    out = {}
    for case in CASES:
        k = synth(*case)
        for keep in keep_modes:
            c, g = run_joint(k, "bf16", keep=keep)
            out[(case, keep)] = (c, g)
        out[(case, "fp32")] = run_joint(k, "fp32")
    return out


def report(tag, a, b):
    (ca, ga), (cb, gb) = a, b
    line = "%-44s cost %.2e" % (tag, np.max(np.abs(ca - cb) / np.maximum(np.abs(cb), 1e-9)))
    worst = 0.0
    for x, y, n in zip(ga, gb, NAMES):
        rel = np.linalg.norm(x.astype(np.float64) - y) / max(np.linalg.norm(y), 1e-30)
        finite = np.isfinite(x).all()
        line += "  %s %.2e%s" % (n, rel, "" if finite else " NONFINITE")
        worst = max(worst, rel if finite else 1e9)
    print(line + ("   <-- BAD" if worst > 5e-2 else ""), flush=True)
    return worst


if __name__ == "__main__":
    res = run_all()
    bad = 0.0
    for case in CASES:
        for keep in (True, False):
            bad = max(bad, report("%s keep=%d vs fp32" % (case[:5], keep), res[(case, keep)], res[(case, "fp32")]))
        (c1, g1), (c2, g2) = res[(case, True)], res[(case, False)]
        report("%s keep vs recompute" % (case[:5],), (c1, g1), (c2, g2))
    print("WORST", bad)
    sys.exit(1 if bad > 5e-2 else 0)
