"""Data-parallel use of the hot path: one process per GPU, utterances sharded across ranks, ONE
collective per step.

Reference being replaced: tf.distribute.MirroredStrategy in run_rnnt.py:119-127 (in-graph replication,
NCCL all-reduce of every gradient inside optimizer.apply_gradients, run_rnnt.py:288, plus
strategy.reduce(MEAN) of the per-example losses, run_rnnt.py:292-296).  For this path the only
cross-replica quantities are the loss sum and the joint's weight gradients (dW, db): they are packed
into one flat fp32 buffer and reduced with a single all-reduce (NCCL over NVLink on GPUs, gloo in the
CPU tests); d_enc / d_pred are per-utterance and never leave the rank.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced shard [lo, hi) of n_items utterances for `rank` (first n%world ranks get one more)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_assignment(work, world_size):
    """Greedy longest-first assignment of utterances to ranks by lattice size T_b*U_b (ragged batches,
    SURVEY 8e).  Returns a list of index lists, one per rank; deterministic."""
    order = sorted(range(len(work)), key=lambda i: (-int(work[i]), i))
    loads = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(work[i])
    return [sorted(ix) for ix in out]


def pack(loss_sum, dW, db):
    return torch.cat([loss_sum.reshape(1).to(torch.float32), dW.reshape(-1), db.reshape(-1)])


def unpack(buf, dW_shape, db_shape):
    nW = 1
    for s in dW_shape:
        nW *= s
    return buf[0], buf[1:1 + nW].view(dW_shape), buf[1 + nW:].view(db_shape)


def allreduce_loss_and_weight_grads(loss_sum, dW, db, group=None):
    """Sum [loss_sum | dW | db] over ranks with one collective; returns (loss_sum, dW, db) views of the
    reduced buffer.  A no-op (besides packing) when torch.distributed is not initialised."""
    buf = pack(loss_sum, dW, db)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack(buf, dW.shape, db.shape)


def allreduce_packed_(tensors, group=None):
    """In-place SUM over ranks of a list of fp32 tensors with ONE collective (flatten -> all_reduce -> copy back).
    The general form of allreduce_loss_and_weight_grads for callers whose joint has more parameters than (W, b)
    (the un-hoisted Dense-1 of model.py:162-163).  A no-op when torch.distributed is not initialised or world == 1."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return tensors
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors
