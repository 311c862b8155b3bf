"""bf16 tensor-core path (tcgen05 fused joint kernels) against the fp64 oracle (small) and the fp32
exact CUDA path (BASELINE C2 size).  bf16 operands carry 2^-9 relative rounding, so this path is
held to its own measured tolerance (DESIGN.md "Tolerances"), not to the fp32 rtol 1e-4 gate."""
import numpy as np
import pytest
import torch

from conftest import assert_close
from test_gpu_joint import run_joint, synth

pytestmark = pytest.mark.gpu

BF16_COST_RTOL = 2e-3
BF16_GRAD_NTOL = 2e-2     # norm-wise: |err| <= ntol * max|grad|


@pytest.mark.parametrize("B,T,U,V,H,seed,ragged", [
    (2, 20, 8, 64, 64, 0, False),        # one 16x8 tile shape, single K block, single V chunk
    (3, 37, 19, 128, 128, 1, True),      # ragged, partially filled tiles
    (2, 50, 40, 192, 320, 2, True),      # V chunk of 64 x3, 5 K blocks
    (1, 9, 140, 64, 64, 3, False),       # U > 128: two u-blocks per time step
    (2, 33, 128, 512, 640, 4, True),     # the BASELINE C3 tile geometry (1x128 tiles, 10 K blocks, NC=256)
])
@pytest.mark.parametrize("keep", [True, False])   # backward from the kept activations / by recomputing the projection
def test_bf16_vs_oracle(oracle, B, T, U, V, H, seed, ragged, keep):
    k = synth(B, T, U, V, H, seed, ragged)
    gs = np.linspace(0.5, 1.5, B)
    o = oracle.joint_loss_grad(*(k[n].astype(np.float64) for n in ("enc", "pred", "W", "b")), k["labels"],
                               k["input_lengths"], k["label_lengths"], 0, grad_scale=gs)
    costs, grads = run_joint(k, "bf16", scale=gs, keep=keep)
    assert_close(costs, o["costs"], rtol=BF16_COST_RTOL, atol=1e-2, what="costs")
    for g, n in zip(grads, ("d_enc", "d_pred", "dW", "db")):
        assert_close(g, o[n], rtol=0, atol=0, ntol=BF16_GRAD_NTOL, what=n)
    # padded positions of d_enc / d_pred are exactly zero
    for b in range(B):
        assert not grads[0][b, k["input_lengths"][b]:].any()
        assert not grads[1][b, k["label_lengths"][b] + 1:].any()


def test_bf16_vs_fp32_at_c2():
    k = synth(16, 256, 64, 256, 320, 5, ragged=True)
    c32, g32 = run_joint(k, "fp32")
    c16, g16 = run_joint(k, "bf16")
    assert_close(c16, c32, rtol=BF16_COST_RTOL, atol=1e-2, what="costs")
    for a, b, n in zip(g16, g32, ("d_enc", "d_pred", "dW", "db")):
        assert_close(a, b, rtol=0, atol=0, ntol=BF16_GRAD_NTOL, what=n)
        rel = np.linalg.norm(a - b) / np.linalg.norm(b)
        assert rel < 1e-2, (n, rel)


def test_bf16_common_voice_shaped_batch():
    """BASELINE C5 in miniature: ragged T_b, U_b (one long and one minimal utterance), V=4096, H=640,
    maxU=200 (8-wide u-tiles) -- bf16 tensor-core path against the fp32 exact path."""
    rng = np.random.default_rng(11)
    B, T, U, V, H = 4, 160, 200, 4096, 640
    k = synth(B, T, U, V, H, 11, ragged=False)
    k["input_lengths"] = np.array([T, 10, 97, 160], np.int32)
    k["label_lengths"] = np.array([U - 1, 9, 120, 30], np.int32)
    for b in range(B):
        k["labels"][b, k["label_lengths"][b]:] = 0
    c32, g32 = run_joint(k, "fp32")
    c16, g16 = run_joint(k, "bf16")
    assert np.all(np.isfinite(c16))
    assert_close(c16, c32, rtol=BF16_COST_RTOL, atol=1e-2, what="costs")
    for a, b_, n in zip(g16, g32, ("d_enc", "d_pred", "dW", "db")):
        assert np.all(np.isfinite(a)), n
        rel = np.linalg.norm(a - b_) / np.linalg.norm(b_)
        assert rel < 1e-2, (n, rel)
    for b in range(B):   # padded positions exactly zero
        assert not g16[0][b, k["input_lengths"][b]:].any() and not g16[1][b, k["label_lengths"][b] + 1:].any()


@pytest.mark.parametrize("B,T,U,V,H,blank", [(3, 9, 1, 64, 64, 0),      # empty transcripts (U == 1)
                                             (3, 1, 5, 64, 128, 0),     # a single encoder frame
                                             (2, 12, 6, 128, 64, 5)])   # blank index != 0
@pytest.mark.parametrize("keep", [True, False])
def test_bf16_edge_lattices(oracle, B, T, U, V, H, blank, keep):
    rng = np.random.default_rng(21)
    k = synth(B, T, max(U, 2), V, H, 21, ragged=False)
    k["pred"] = k["pred"][:, :U].copy()
    cand = np.array([v for v in range(V) if v != blank])
    k["labels"] = rng.choice(cand, size=(B, U - 1)).astype(np.int32) if U > 1 else np.zeros((B, 0), np.int32)
    k["label_lengths"] = np.full(B, U - 1, np.int32)
    k["blank"] = np.int32(blank)
    o = oracle.joint_loss_grad(*(k[n].astype(np.float64) for n in ("enc", "pred", "W", "b")), k["labels"],
                               k["input_lengths"], k["label_lengths"], blank, grad_scale=np.full(B, 1.0 / B))
    costs, grads = run_joint(k, "bf16", keep=keep)
    assert_close(costs, o["costs"], rtol=BF16_COST_RTOL, atol=1e-2, what="costs")
    for g, n in zip(grads, ("d_enc", "d_pred", "dW", "db")):
        assert_close(g, o[n], rtol=0, atol=0, ntol=BF16_GRAD_NTOL, what=n)


def test_compacted_and_padded_backward_agree():
    """Ragged batch: the backward over valid tiles only (allow_host_sync) and the sync-free backward over the padded
    tile set produce the same gradients (identical arithmetic per row; only the GEMM reduction order may differ)."""
    k = synth(5, 70, 45, 256, 192, 31, ragged=True)
    for keep in (True, False):
        c1, g1 = run_joint(k, "bf16", compact=True, keep=keep)
        c2, g2 = run_joint(k, "bf16", compact=False, keep=keep)
        assert np.array_equal(c1, c2)
        for a, b_, n in zip(g1, g2, ("d_enc", "d_pred", "dW", "db")):
            assert_close(a, b_, rtol=1e-5, atol=0, ntol=1e-5, what=n)


def test_kept_and_recomputed_backward_agree():
    """keep_activations: the forward leaves fp16 softmax numerators (2^-11 relative) + bf16 tanh outputs and the backward
    is a streaming pass; without it the backward recomputes the projection.  Same operands, so the two differ only by
    the fp16 rounding of the numerators -- an order of magnitude inside the bf16 path's own tolerance."""
    k = synth(6, 90, 50, 320, 256, 41, ragged=True)
    c1, g1 = run_joint(k, "bf16", keep=True)
    c2, g2 = run_joint(k, "bf16", keep=False)
    assert_close(c1, c2, rtol=1e-6, atol=1e-4, what="costs")
    for a, b_, n in zip(g1, g2, ("d_enc", "d_pred", "dW", "db")):
        assert_close(a, b_, rtol=0, atol=0, ntol=2e-3, what=n)
        assert np.linalg.norm(a - b_) <= 2e-3 * np.linalg.norm(b_), n
