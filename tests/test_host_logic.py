"""Host-side logic of the reference-facing Python surface (no GPU needed)."""
import numpy as np
import pytest
import torch

import rnnt_speech_recognition_b200 as rb


def i32(*v):
    return torch.tensor(v, dtype=torch.int32)


def test_encoder_lengths_is_ceil_of_true_division():
    # utils/loss.py:31-33 -- tf.math.ceil(spec_lengths / reduction_factor) -> int32
    spec = torch.tensor([1, 2, 3, 4, 5, 1023, 1024, 1025])
    assert rb.encoder_lengths(spec, 2).tolist() == [1, 1, 2, 2, 3, 512, 512, 513]
    assert rb.encoder_lengths(spec, 1).tolist() == spec.tolist()
    assert rb.encoder_lengths(spec, 3).dtype == torch.int32


def test_certify_inputs_messages():
    acts = torch.zeros(2, 4, 3, 5)
    lab, tl, ll = torch.zeros(2, 2, dtype=torch.int32), i32(4, 4), i32(2, 2)
    rb.certify_inputs(acts, lab, tl, ll)
    with pytest.raises(TypeError, match="labels must be"):
        rb.certify_inputs(acts, lab.long(), tl, ll)
    with pytest.raises(TypeError, match="lengths must be"):
        rb.certify_inputs(acts, lab, tl.long(), ll)
    with pytest.raises(ValueError, match="must be contiguous"):
        rb.certify_inputs(acts.transpose(1, 2), lab, tl, ll)
    with pytest.raises(ValueError, match="must have a length per example"):
        rb.certify_inputs(acts, lab, i32(4), ll)
    with pytest.raises(ValueError, match="log_probs must be 4D"):
        rb.certify_inputs(acts[0], lab, i32(4, 4, 4, 4), i32(2, 2, 2, 2))
    with pytest.raises(ValueError, match="Input length mismatch"):
        rb.certify_inputs(acts, lab, i32(3, 3), ll)
    with pytest.raises(ValueError, match="Output length mismatch"):
        rb.certify_inputs(acts, lab, tl, i32(1, 1))


def test_no_cpu_fallback():
    acts = torch.zeros(1, 2, 3, 5, requires_grad=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rb.rnnt_loss(acts, torch.zeros(1, 2, dtype=torch.int32), i32(2), i32(2))
    loss_fn = rb.get_loss_fn(2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        loss_fn(torch.zeros(1, 2), acts, torch.tensor([4]), torch.tensor([2]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        rb.joint_rnnt_loss(torch.zeros(1, 2, 8), torch.zeros(1, 3, 8), torch.zeros(8, 5), torch.zeros(5),
                           torch.zeros(1, 2, dtype=torch.int32), i32(2), i32(2))


def test_joint_module_parameters_follow_keras_layout():
    j = rb.Joint(proj_size=12, joint_net_size=16, vocab_size=10)
    assert j.kernel_1.shape == (12, 16) and j.kernel_2.shape == (16, 10)      # Dense kernel is (in, out)
    f, g = torch.randn(2, 5, 12), torch.randn(2, 3, 12)
    # the hoisting identity `Joint.hoist` relies on (Dense-1 is linear in front of its tanh); the projections themselves run
    # in the extension (rnntb200_dense1_*), so on a machine without a GPU `hoist` must refuse, not fall back
    e, p = f @ j.kernel_1 + j.bias_1, g @ j.kernel_1
    want = torch.tanh((f[:, :, None] + g[:, None]) @ j.kernel_1 + j.bias_1)   # model.py:158-163 literally
    got = torch.tanh(e[:, :, None] + p[:, None])
    assert torch.allclose(got, want, atol=1e-5)
    with pytest.raises((TypeError, OSError, RuntimeError)):
        j.hoist(f, g)


def test_valid_tile_count_and_the_workspace_it_buys():
    """`valid_tile_count` is the host-side sum the descriptor's `valid_tile_bound` expects, and the workspace size is a pure
    function of the descriptor (no GPU needed): for BASELINE C5's shape the padded numerators need several chunks of 16 GiB,
    the promise of the valid tiles turns that into one chunk sized for them."""
    import ctypes as C
    from rnnt_speech_recognition_b200 import _lib
    assert rb.valid_tile_count([16, 17, 1], [7, 8, 0]) == 1 * 1 + 2 * 2 + 1 * 1       # ceil(T/16) * ceil((U+1)/8)
    L = _lib.load(build_if_missing=True)
    B, T, U, V, H = 64, 1600, 200, 4096, 640
    rng = np.random.default_rng(0)
    il, ll = rng.integers(100, T + 1, B), rng.integers(10, U + 1, B) - 1
    vt = rb.valid_tile_count(il, ll)
    sizes = []
    for bound in (0, vt):
        d = _lib.JointDesc(B, T, U, H, V, 0, 1, None, bound, 1)
        s = C.c_size_t(0)
        assert L.rnntb200_joint_workspace_size(C.byref(d), C.byref(s)) == 0
        sizes.append(s.value)
    per_row = V * 2 + 8
    assert sizes[0] < 40 << 30                                    # padded: chunks of <= 16 GiB of numerators (+ planes)
    assert sizes[1] >= vt * 128 * per_row                         # one chunk that holds every valid row block
    assert sizes[1] < 64 * ((T + 15) // 16) * ((U + 7) // 8) * 128 * per_row / 2     # far below the padded lattice
