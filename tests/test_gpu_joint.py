"""GPU parity of the fused joint + loss path (rnntb200_joint_loss_forward/backward through the
torch surface) against golden fixtures (torch-fp64 + reference library) and the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import assert_close, golden_names, load

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def run_joint(k, precision, scale=None, compact=True, keep=None, valid_tiles=None):
    import rnnt_speech_recognition_b200 as rb
    t = [dev(k[n], torch.float32).requires_grad_() for n in ("enc", "pred", "W", "b")]
    lab, il, ll = (dev(k[n], torch.int32) for n in ("labels", "input_lengths", "label_lengths"))
    costs = rb.joint_rnnt_loss(*t, lab, il, ll, blank=int(k["blank"]), precision=precision, compact=compact,
                               keep_activations=keep, valid_tiles=valid_tiles)
    B = costs.shape[0]
    w = torch.full((B,), 1.0 / B, device="cuda") if scale is None else dev(scale, torch.float32)
    (costs * w).sum().backward()                    # run_rnnt.py:278: sum / global batch
    return costs.detach().cpu().numpy(), [x.grad.cpu().numpy() for x in t]


@pytest.mark.parametrize("name", golden_names("joint_"))
def test_golden_joint_fp32(name):
    k = load(name)                                   # BASELINE C1 is joint_c1_B2_T20_U8_V32_H64
    costs, grads = run_joint(k, "fp32")
    assert_close(costs, k["costs"], rtol=1e-5, atol=1e-5, what="costs")
    for g, n in zip(grads, ("d_enc", "d_pred", "dW", "db")):
        assert_close(g, k[n], rtol=1e-4, atol=1e-6, ntol=1e-5, what=n)


def synth(B, T, U, V, H, seed, ragged=True):
    rng = np.random.default_rng(seed)
    k = dict(enc=rng.standard_normal((B, T, H)).astype(np.float32), pred=rng.standard_normal((B, U, H)).astype(np.float32),
             W=(rng.standard_normal((H, V)) / np.sqrt(H)).astype(np.float32), b=(0.1 * rng.standard_normal(V)).astype(np.float32),
             labels=rng.integers(1, V, (B, U - 1)).astype(np.int32), blank=np.int32(0))
    il, ll = np.full(B, T, np.int32), np.full(B, U - 1, np.int32)
    if ragged:
        il, ll = rng.integers(1, T + 1, B).astype(np.int32), rng.integers(0, U, B).astype(np.int32)
        il[0], ll[0] = T, U - 1
    k.update(input_lengths=il, label_lengths=ll)
    return k


@pytest.mark.parametrize("B,T,U,V,H,seed", [(3, 37, 19, 72, 96, 0), (2, 50, 40, 130, 70, 1)])
def test_fp32_vs_oracle(oracle, B, T, U, V, H, seed):
    k = synth(B, T, U, V, H, seed)
    gs = np.linspace(0.5, 1.5, B)
    o = oracle.joint_loss_grad(*(k[n].astype(np.float64) for n in ("enc", "pred", "W", "b")), k["labels"],
                               k["input_lengths"], k["label_lengths"], 0, grad_scale=gs)
    costs, grads = run_joint(k, "fp32", scale=gs)
    assert_close(costs, o["costs"], rtol=1e-5, atol=1e-4, what="costs")
    for g, n in zip(grads, ("d_enc", "d_pred", "dW", "db")):
        assert_close(g, o[n], rtol=1e-4, atol=1e-6, ntol=1e-5, what=n)


def test_fused_equals_op_on_materialised_logits():
    """Size-independent property at BASELINE C2 size (B=16,T=256,U=64,V=256,H=320): the fused fp32 path
    and the reference-shaped two-step path (Joint.forward -> rnnt_loss, run_rnnt.py:269-273) agree on
    costs and on all four input gradients (autograd chains the second one through torch ops)."""
    import rnnt_speech_recognition_b200 as rb
    k = synth(16, 256, 64, 256, 320, 5, ragged=True)
    costs, grads = run_joint(k, "fp32")
    t = [dev(k[n], torch.float32).requires_grad_() for n in ("enc", "pred", "W", "b")]
    lab, il, ll = (dev(k[n], torch.int32) for n in ("labels", "input_lengths", "label_lengths"))
    logits = rb.joint_logits(*[x.detach() for x in t])
    ref_logits = torch.tanh(t[0][:, :, None] + t[1][:, None]) @ t[2] + t[3]
    assert torch.allclose(logits, ref_logits.detach(), atol=2e-5, rtol=1e-5)
    c2 = rb.rnnt_loss(ref_logits, lab, il, ll)
    (c2.sum() / 16).backward()
    assert_close(costs, c2.detach().cpu().numpy(), rtol=1e-5, atol=1e-3, what="costs")
    for g, x, n in zip(grads, t, ("d_enc", "d_pred", "dW", "db")):
        assert_close(g, x.grad.cpu().numpy(), rtol=1e-4, atol=1e-6, ntol=2e-5, what=n)
    assert abs(grads[3].sum()) < 1e-3 * np.abs(grads[3]).sum() + 1e-5     # sum_v db == 0 (rows of dlogits sum to 0)


def test_joint_module_forward_and_decode_step():
    """Joint.forward == model.py:158-166 literally (un-hoisted Dense-1), Joint.step == utils/decoding.py:6-18."""
    import rnnt_speech_recognition_b200 as rb
    torch.manual_seed(3)
    j = rb.Joint(proj_size=48, joint_net_size=64, vocab_size=40, precision="fp32").cuda()
    f, g = torch.randn(2, 7, 48, device="cuda"), torch.randn(2, 5, 48, device="cuda")
    with torch.no_grad():
        want = torch.tanh((f[:, :, None] + g[:, None]) @ j.kernel_1 + j.bias_1) @ j.kernel_2 + j.bias_2
        got = j(f, g)
        assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)
        step = j.step(f, g)
        want_step = (torch.tanh((f[:, :, None] + g[:, -1:, :][:, None]) @ j.kernel_1 + j.bias_1) @ j.kernel_2 + j.bias_2)[:, 0, 0, :]
        assert step.shape == (2, 40) and torch.allclose(step, want_step, atol=2e-5, rtol=1e-5)
    # fused loss through the module: gradients reach the Keras-layout parameters of both Dense layers
    lab = torch.randint(1, 40, (2, 4), dtype=torch.int32, device="cuda")
    il, ll = torch.tensor([7, 5], dtype=torch.int32, device="cuda"), torch.tensor([4, 2], dtype=torch.int32, device="cuda")
    costs = j.loss(f, g, lab, il, ll)
    costs.sum().backward()
    ref = rb.rnnt_loss(torch.tanh((f[:, :, None] + g[:, None]) @ j.kernel_1 + j.bias_1) @ j.kernel_2 + j.bias_2, lab, il, ll)
    assert torch.allclose(costs, ref, rtol=1e-5, atol=1e-4)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in j.parameters())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_step_runs_and_learns(precision):
    """run_rnnt.py:259-296 for the joint -- loss / global batch, backward, (packed all-reduce: world 1
    here, world 2 in tests/test_dist_gloo.py), SGD-momentum step.  A few steps on a fixed batch must lower the loss,
    and the encoder / prediction inputs must receive gradients for the caller's networks."""
    import rnnt_speech_recognition_b200 as rb
    torch.manual_seed(3)
    B, T, U, P, H, V = 3, 14, 5, 64, 64, 64
    joint = rb.Joint(P, H, V, precision=precision).cuda()
    opt = rb.make_optimizer(joint.parameters(), learning_rate=0.02)
    enc = torch.randn(B, T, P, device="cuda", requires_grad=True)
    pred = torch.randn(B, U, P, device="cuda", requires_grad=True)
    labels = torch.randint(1, V, (B, U - 1), device="cuda", dtype=torch.int32)
    spec_lengths = torch.tensor([2 * T, 2 * T - 3, 2 * T - 1], device="cuda", dtype=torch.int32)   # ceil(./2) -> T, T-1, T
    label_lengths = torch.tensor([U - 1, U - 2, U - 1], device="cuda", dtype=torch.int32)
    losses = [rb.joint_train_step(joint, opt, enc, pred, labels, spec_lengths, label_lengths, global_batch_size=B).item()
              for _ in range(6)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses
    assert enc.grad is not None and torch.isfinite(enc.grad).all() and enc.grad.abs().sum() > 0
    assert pred.grad is not None and torch.isfinite(pred.grad).all()
    assert not enc.grad[1, T - 1:].any()          # frames past ceil(spec_length / 2) receive no gradient
