"""World-size-2 gloo test of the multi-rank host logic (sharding + the single packed all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rnnt_speech_recognition_b200 import distributed as D


def test_shard_bounds_cover_everything():
    for n in (1, 7, 32, 33, 64):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = D.shard_bounds(n, w, r)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_balanced_assignment():
    work = [1600 * 200, 100 * 10, 800 * 100, 800 * 100, 400 * 50, 1200 * 150]
    parts = D.balanced_assignment(work, 2)
    assert sorted(sum(parts, [])) == list(range(len(work)))
    loads = [sum(work[i] for i in p) for p in parts]
    assert max(loads) / sum(loads) < 0.62


def _worker(rank, world, port, H, V, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    loss, dW, db = torch.rand(1, generator=g)[0], torch.rand(H, V, generator=g), torch.rand(V, generator=g)
    l2, w2, b2 = D.allreduce_loss_and_weight_grads(loss, dW, db)
    out[rank] = (l2.item(), w2.clone().numpy(), b2.clone().numpy())
    dist.destroy_process_group()


def test_packed_allreduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    H, V, world = 6, 5, 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, H, V, out), nprocs=world, join=True)
        res = dict(out)
    want_l, want_w, want_b = 0.0, np.zeros((H, V), np.float32), np.zeros(V, np.float32)
    for r in range(world):
        g = torch.Generator().manual_seed(100 + r)
        want_l += torch.rand(1, generator=g)[0].item()
        want_w += torch.rand(H, V, generator=g).numpy()
        want_b += torch.rand(V, generator=g).numpy()
    for r in range(world):
        l, w, b = res[r]
        assert abs(l - want_l) < 1e-6 and np.allclose(w, want_w) and np.allclose(b, want_b)


class _StubJoint(torch.nn.Module):
    """Pure-torch stand-in with the Joint.loss signature (the fused loss itself needs a GPU): per-utterance costs that
    depend on every parameter, so the train-step plumbing (global-batch scaling, packed all-reduce, SGD) can be checked."""

    def __init__(self, P, V):
        super().__init__()
        g = torch.Generator().manual_seed(7)
        self.kernel_2 = torch.nn.Parameter(torch.randn(P, V, generator=g) * 0.3)
        self.bias_2 = torch.nn.Parameter(torch.randn(V, generator=g) * 0.1)

    def loss(self, inp_enc, pred_outputs, labels, input_lengths, label_lengths):
        z = torch.tanh(inp_enc.mean(1) + pred_outputs.mean(1))
        return (((z @ self.kernel_2 + self.bias_2) ** 2).sum(-1) + 0.01 * input_lengths.float() + 0.0 * labels.float().sum(-1)
                + 0.0 * label_lengths.float())


def _batch(B, T, U, P):
    g = torch.Generator().manual_seed(5)
    return (torch.randn(B, T, P, generator=g), torch.randn(B, U, P, generator=g), torch.randint(1, 9, (B, U - 1), generator=g),
            torch.full((B,), 2 * T), torch.full((B,), U - 1))


def _train_worker(rank, world, port, out):
    from rnnt_speech_recognition_b200 import joint_train_step, make_optimizer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, T, U, P, V = 4, 6, 3, 8, 5
    enc, pred, lab, sl, ll = _batch(B, T, U, P)
    lo, hi = D.shard_bounds(B, world, rank)
    joint = _StubJoint(P, V)
    opt = make_optimizer(joint.parameters(), learning_rate=0.1)
    losses = [joint_train_step(joint, opt, enc[lo:hi], pred[lo:hi], lab[lo:hi], sl[lo:hi], ll[lo:hi], global_batch_size=B).item()
              for _ in range(3)]
    out[rank] = (losses, joint.kernel_2.detach().clone().numpy(), joint.bias_2.detach().clone().numpy())
    dist.destroy_process_group()


def test_train_step_world2_matches_single_process():
    """run_rnnt.py:278-296 semantics: sum of per-rank (costs.sum()/global_batch) gradients == full-batch gradient; every
    rank applies the same update and reports the global mean loss."""
    from rnnt_speech_recognition_b200 import joint_train_step, make_optimizer
    B, T, U, P, V = 4, 6, 3, 8, 5
    enc, pred, lab, sl, ll = _batch(B, T, U, P)
    ref = _StubJoint(P, V)
    opt = make_optimizer(ref.parameters(), learning_rate=0.1)
    ref_losses = [joint_train_step(ref, opt, enc, pred, lab, sl, ll, global_batch_size=B).item() for _ in range(3)]
    assert ref_losses[2] < ref_losses[0]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_train_worker, args=(2, port, out), nprocs=2, join=True)
        res = dict(out)
    for r in range(2):
        losses, k2, b2 = res[r]
        assert np.allclose(losses, ref_losses, rtol=1e-5, atol=1e-6)
        assert np.allclose(k2, ref.kernel_2.detach().numpy(), rtol=1e-5, atol=1e-6)
        assert np.allclose(b2, ref.bias_2.detach().numpy(), rtol=1e-5, atol=1e-6)
