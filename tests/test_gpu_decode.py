"""the greedy-decode joint (reference utils/decoding.py:6-18 + the log_softmax / argmax of decoding.py:69-78)
as one launch of `rnntb200_joint_step`, against the float64 numpy restatement in oracle/oracle.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_ATOL = 2e-5      # fp32 FMA chains of length P and H against float64


def _case(B, P, H, V, seed, scale=1.0):
    g = np.random.default_rng(seed)
    f = g.standard_normal((B, P)).astype(np.float32)
    q = g.standard_normal((B, P)).astype(np.float32)
    K1 = (g.standard_normal((P, H)) / np.sqrt(P)).astype(np.float32)
    b1 = (0.1 * g.standard_normal(H)).astype(np.float32)
    K2 = (scale * g.standard_normal((H, V)) / np.sqrt(H)).astype(np.float32)
    b2 = (0.1 * g.standard_normal(V)).astype(np.float32)
    return f, q, K1, b1, K2, b2


@pytest.mark.parametrize("B,P,H,V", [(1, 640, 640, 1024),       # the reference's decoder: one utterance, hparams-sized joint
                                     (1, 320, 512, 4096),
                                     (3, 77, 50, 33),            # nothing a multiple of anything
                                     (2, 8, 5, 3),               # fewer hidden units / vocabulary entries than CTAs in the cluster
                                     (5, 1024, 1024, 256),
                                     (16, 128, 96, 1000)])
def test_joint_step_vs_oracle(B, P, H, V):
    import rnnt_speech_recognition_b200 as rb
    from oracle import oracle
    f, q, K1, b1, K2, b2 = _case(B, P, H, V, seed=B * 1000 + V)
    t = [torch.from_numpy(a).cuda() for a in (f, q, K1, b1, K2, b2)]
    logits, best, logp = rb.joint_step(*t, want_logits=True, want_best=True)
    y, ybest, ylogp = oracle.joint_step(f, q, K1, b1, K2, b2)
    np.testing.assert_allclose(logits.cpu().numpy(), y, atol=LOGIT_ATOL * max(1.0, np.abs(y).max()), rtol=0)
    # the argmax must be the reference's wherever float32 can tell the two best apart
    srt = np.sort(y, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-4 if V > 1 else np.ones(B, bool)
    assert clear.any()
    np.testing.assert_array_equal(best.cpu().numpy()[clear], ybest[clear])
    np.testing.assert_allclose(logp.cpu().numpy()[clear], ylogp[clear], atol=5e-5, rtol=0)
    # best-only call: same answers without the (B,V) store
    _, best2, logp2 = rb.joint_step(*t, want_logits=False, want_best=True)
    assert torch.equal(best2, best) and torch.equal(logp2, logp)


def test_joint_step_projected_inputs_and_strided_views():
    """K1 = None (already projected activations) and row-strided views (`encoded[:, i, :]`, `pred_out[:, -1, :]`)."""
    import rnnt_speech_recognition_b200 as rb
    from oracle import oracle
    g = np.random.default_rng(5)
    enc = g.standard_normal((2, 7, 192)).astype(np.float32)
    pred = g.standard_normal((2, 4, 192)).astype(np.float32)
    K2 = (g.standard_normal((192, 300)) / 14).astype(np.float32)
    b2 = g.standard_normal(300).astype(np.float32)
    te, tp, tk, tb = (torch.from_numpy(a).cuda() for a in (enc, pred, K2, b2))
    logits, best, _ = rb.joint_step(te[:, 3, :], tp[:, -1, :], None, None, tk, tb, want_logits=True, want_best=True)
    y, ybest, _ = oracle.joint_step(enc[:, 3, :], pred[:, -1, :], None, None, K2, b2)
    np.testing.assert_allclose(logits.cpu().numpy(), y, atol=LOGIT_ATOL * np.abs(y).max(), rtol=0)
    np.testing.assert_array_equal(best.cpu().numpy(), ybest)


def test_joint_module_step_matches_lattice_forward():
    """`Joint.step(f, g)` is `joint(model, f, g)` of decoding.py: frame 0 of f against the LAST frame of g -- i.e. cell
    (0, U-1) of the lattice the training-time forward produces; `greedy_step` is its argmax."""
    import rnnt_speech_recognition_b200 as rb
    torch.manual_seed(3)
    j = rb.Joint(96, 128, 200, precision="fp32").cuda()
    f = torch.randn(2, 5, 96, device="cuda")
    gq = torch.randn(2, 3, 96, device="cuda")
    with torch.no_grad():
        e, q = j.hoist(f, gq)                    # Dense-1 on the un-broadcast inputs, in the extension
        want = torch.tanh((f[:, :, None] + gq[:, None]).double() @ j.kernel_1.double() + j.bias_1.double())   # model.py:158-163 literally
        torch.testing.assert_close(torch.tanh(e[:, :, None] + q[:, None]).double(), want, atol=1e-5, rtol=0)
        lattice = j(f, gq)                       # (B,T,U,V) through the training-time kernels
        step = j.step(f, gq)
        best, logp = j.greedy_step(f, gq)
    torch.testing.assert_close(step, lattice[:, 0, -1, :], atol=2e-5, rtol=1e-5)
    ref = torch.log_softmax(step.double(), dim=-1)
    assert torch.equal(best.long(), ref.argmax(dim=-1))
    torch.testing.assert_close(logp.double(), ref.max(dim=-1).values, atol=5e-5, rtol=0)


def test_joint_step_rejects_cpu_tensors():
    import rnnt_speech_recognition_b200 as rb
    with pytest.raises(TypeError):
        rb.joint_step(torch.zeros(1, 8), torch.zeros(1, 8), None, None, torch.zeros(8, 4), None)


# ---------------------------------------------------------------------------------------------------------------
# Dense-1 (model.py:162-163) in the extension, forward and the three gradients
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,P,H,bias", [((4, 50), 640, 640, True), ((3, 7), 80, 64, False), ((2, 33), 77, 50, True),
                                           ((32 * 512,), 640, 640, True)])
def test_dense1_forward_backward_vs_float64(shape, P, H, bias):
    import rnnt_speech_recognition_b200 as rb
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*shape, P, generator=g).cuda().requires_grad_()
    K = (torch.randn(P, H, generator=g) / P ** 0.5).cuda().requires_grad_()
    b = (0.1 * torch.randn(H, generator=g)).cuda().requires_grad_() if bias else None
    up = torch.randn(*shape, H, generator=g).cuda()
    out = rb.dense1(x, K, b)
    out.backward(up)
    xd, Kd, ud = x.detach().double(), K.detach().double(), up.double()
    ref = xd @ Kd + (b.detach().double() if bias else 0.0)
    rows = xd.reshape(-1, P)
    scale = lambda t: float(t.abs().max())
    assert float((out.detach().double() - ref).abs().max()) <= 2e-6 * max(1.0, scale(ref)) * P ** 0.5
    r_dx, r_dk = ud @ Kd.T, rows.T @ ud.reshape(-1, H)
    assert float((x.grad.double() - r_dx).abs().max()) <= 2e-6 * scale(r_dx) * H ** 0.5
    assert float((K.grad.double() - r_dk).abs().max()) <= 2e-6 * scale(r_dk) * rows.shape[0] ** 0.5
    if bias:
        r_db = ud.reshape(-1, H).sum(0)
        assert float((b.grad.double() - r_db).abs().max()) <= 2e-6 * max(1.0, scale(r_db)) * rows.shape[0] ** 0.5


def test_joint_module_full_joint_gradients_reach_dense1():
    """model.py's whole joint (P -> H tanh -> V) through the extension: Joint.loss backpropagates into kernel_1 / bias_1 and
    the un-projected inputs, and agrees with the same computation through torch eager float64 + the oracle-checked loss."""
    import rnnt_speech_recognition_b200 as rb
    torch.manual_seed(9)
    B, T, U, P, H, V = 2, 12, 5, 48, 64, 40
    j = rb.Joint(P, H, V, precision="fp32").cuda()
    f = torch.randn(B, T, P, device="cuda", requires_grad=True)
    q = torch.randn(B, U, P, device="cuda", requires_grad=True)
    lab = torch.randint(1, V, (B, U - 1), dtype=torch.int32, device="cuda")
    il = torch.tensor([T, T - 3], dtype=torch.int32, device="cuda")
    ll = torch.tensor([U - 1, U - 2], dtype=torch.int32, device="cuda")
    (j.loss(f, q, lab, il, ll).sum()).backward()
    got = [t.grad.clone() for t in (f, q, j.kernel_1, j.bias_1, j.kernel_2, j.bias_2)]
    for t in (f, q, j.kernel_1, j.bias_1, j.kernel_2, j.bias_2):
        t.grad = None
    # the same graph with torch eager projections in float64 in front of the (materialised-logits) loss op of this library
    ea = (f.double() @ j.kernel_1.double() + j.bias_1.double()).float()
    pa = (q.double() @ j.kernel_1.double()).float()
    rb.joint_rnnt_loss(ea, pa, j.kernel_2, j.bias_2, lab, il, ll, precision="fp32").sum().backward()
    want = [t.grad for t in (f, q, j.kernel_1, j.bias_1, j.kernel_2, j.bias_2)]
    for a, b in zip(got, want):
        torch.testing.assert_close(a, b, atol=2e-5 * float(b.abs().max()) + 1e-7, rtol=1e-4)
