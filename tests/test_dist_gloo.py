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
