"""RNN-T loss op on materialised logits -- the reference's op-level surface on torch tensors.

Mirrors, name for name:
  * ``warprnnt_tensorflow.rnnt_loss(acts, labels, input_lengths, label_lengths, blank_label=0)``
    (warp-transducer/tensorflow_binding/warprnnt_tensorflow/__init__.py:9-42): per-utterance costs
    ``(B,)``, differentiable w.r.t. ``acts`` with gradient ``grad_loss[:,None,None,None] * grads``.
  * ``warprnnt_pytorch.rnnt_loss / RNNTLoss`` with reductions and ``certify_inputs``
    (warp-transducer/pytorch_binding/warprnnt_pytorch/__init__.py:10-140) -> ``torch_rnnt_loss``,
    ``RNNTLoss``.

Differences from the reference bindings, all deliberate:
  * costs stay on the device (the reference copies them to host and synchronises the stream
    every call, gpu_rnnt.h:209-213; the torch binding builds a CPU tensor, __init__.py:27);
  * CUDA tensors only: the reference's CPU branch is not reproduced (no CPU fallback);
  * gradients of padded cells are written by the gradient kernel itself (no memset pass).
"""
import ctypes as C

import torch

from . import _lib

__all__ = ["rnnt_loss", "torch_rnnt_loss", "RNNTLoss", "certify_inputs"]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check_type(var, t, name):
    if var.dtype is not t:
        raise TypeError("{} must be {}".format(name, t))


def check_contiguous(var, name):
    if not var.is_contiguous():
        raise ValueError("{} must be contiguous".format(name))


def check_dim(var, dim, name):
    if len(var.shape) != dim:
        raise ValueError("{} must be {}D".format(name, dim))


def certify_inputs(log_probs, labels, lengths, label_lengths, check_max=True):
    """Same checks, messages and exception types as warprnnt_pytorch/__init__.py:115-140.
    ``check_max=False`` skips the two device-synchronising max() comparisons (used by the
    TF-flavoured entry, whose op only checks ranks: warprnnt_op.cc:51-86)."""
    check_type(labels, torch.int32, "labels")
    check_type(label_lengths, torch.int32, "label_lengths")
    check_type(lengths, torch.int32, "lengths")
    check_contiguous(log_probs, "log_probs")
    check_contiguous(labels, "labels")
    check_contiguous(label_lengths, "label_lengths")
    check_contiguous(lengths, "lengths")
    if lengths.shape[0] != log_probs.shape[0]:
        raise ValueError("must have a length per example.")
    if label_lengths.shape[0] != log_probs.shape[0]:
        raise ValueError("must have a label length per example.")
    check_dim(log_probs, 4, "log_probs")
    check_dim(labels, 2, "labels")
    check_dim(lengths, 1, "lenghts")
    check_dim(label_lengths, 1, "label_lenghts")
    if check_max:
        max_T = torch.max(lengths)
        max_U = torch.max(label_lengths)
        T, U = log_probs.shape[1:3]
        if T != max_T:
            raise ValueError("Input length mismatch")
        if U != max_U + 1:
            raise ValueError("Output length mismatch")


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: rnnt_b200 has no CPU path (no fallback by design)" % name)


def workspace_bytes(maxT, maxU, B, dtype_size=4):
    sz = C.c_size_t(0)
    _lib.check(_lib.load().get_workspace_size(maxT, maxU, B, True, C.byref(sz), dtype_size), "get_workspace_size")
    return sz.value


class _RNNTOp(torch.autograd.Function):
    """WarpRNNT op + _RNNTLossGrad: gradients are computed eagerly in forward and cached, exactly
    as both reference bindings do (warprnnt_op.cc:100-137, warprnnt_pytorch/__init__.py:24-42)."""

    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, blank):
        L = _lib.load()
        for t, n in ((acts, "acts"), (labels, "labels"), (act_lens, "input_lengths"), (label_lens, "label_lengths")):
            _require_cuda(t, n)
        if acts.dtype != torch.float32:
            raise TypeError("acts must be torch.float32")      # op registered float32-only, warprnnt_op.cc:15
        B, T, U, V = acts.shape
        if labels.shape[1] != U - 1 and not (U == 1 and labels.shape[1] == 0):
            raise ValueError("labels must be (B, U-1)")
        need_grad = acts.requires_grad
        grads = torch.empty_like(acts) if need_grad else None
        costs = torch.empty(B, dtype=torch.float32, device=acts.device)
        ws = torch.empty(workspace_bytes(T, U, B), dtype=torch.uint8, device=acts.device)
        lab = labels if labels.numel() else torch.zeros(1, dtype=torch.int32, device=acts.device)
        with torch.cuda.device(acts.device):     # the stream must be the current stream OF THE TENSORS' DEVICE
            opt = _lib.RnntOptions(_lib.RNNT_GPU, 0, _stream().value, int(blank), T, U, True)
            st = L.rnntb200_loss_device(_ptr(acts), _ptr(grads), _ptr(lab), _ptr(label_lens), _ptr(act_lens), None,
                                        V, B, _ptr(costs), _ptr(ws), opt)
        _lib.check(st, "rnntb200_loss_device")
        ctx.grads = grads
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        g = ctx.grads
        ctx.grads = None
        return g.mul_(grad_costs.view(-1, 1, 1, 1).to(g)), None, None, None, None


def rnnt_loss(acts, labels, input_lengths, label_lengths, blank_label=0):
    """TF-binding signature (warprnnt_tensorflow/__init__.py:9): per-utterance costs (B,).
    ``acts`` are raw logits (B,T,U,V) float32 -- the softmax is performed inside, as the reference's
    GPU op does.  ``labels`` (B,U-1) int32 zero padded; lengths (B,) int32."""
    acts = acts.contiguous()
    certify_inputs(acts, labels, input_lengths, label_lengths, check_max=False)
    return _RNNTOp.apply(acts, labels, input_lengths, label_lengths, blank_label)


def torch_rnnt_loss(acts, labels, act_lens, label_lens, blank=0, reduction="mean"):
    """warprnnt_pytorch.rnnt_loss (pytorch_binding/warprnnt_pytorch/__init__.py:53-70):
    reduction in {'none','sum','mean'}; 'mean' divides the summed cost by the batch size."""
    certify_inputs(acts, labels, act_lens, label_lens, check_max=True)
    costs = _RNNTOp.apply(acts, labels, act_lens, label_lens, blank)
    if reduction in ("sum", "mean"):
        costs = costs.sum().unsqueeze(-1)
        if reduction == "mean":
            costs = costs / acts.size(0)
    elif reduction != "none":
        raise ValueError("reduction must be 'none', 'sum' or 'mean'")
    return costs


class RNNTLoss(torch.nn.Module):
    """warprnnt_pytorch.RNNTLoss (pytorch_binding/warprnnt_pytorch/__init__.py:73-100)."""

    def __init__(self, blank=0, reduction="mean"):
        super().__init__()
        self.blank = blank
        self.reduction = reduction

    def forward(self, acts, labels, act_lens, label_lens):
        return torch_rnnt_loss(acts, labels, act_lens, label_lens, self.blank, self.reduction)
