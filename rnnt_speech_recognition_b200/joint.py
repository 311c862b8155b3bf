"""Fused joint network + RNN-T loss (the north-star path): torch surface over
``rnntb200_joint_loss_forward/backward`` (include/rnnt_b200.h).

Reference being replaced (file:line under /root/reference):
  model.py:158-166      joint_inp = enc[:,:,None,:] + pred[:,None,:,:]; Dense(tanh); Dense(vocab)
  utils/loss.py:24-36   loss adapter -> warprnnt_tensorflow.rnnt_loss
  run_rnnt.py:269-284   model(...) -> loss_fn(...) -> tape.gradient

``enc_acts`` / ``pred_acts`` are the Dense-1 projections hoisted out of the lattice
(W1^T(f_t+g_u)+b1 = (W1^T f_t + b1) + W1^T g_u): the (B,T,U,*) tensors of the
reference never exist here, in either direction.
"""
import ctypes as C
import os

import torch

from . import _lib
from .loss import encoder_lengths

__all__ = ["joint_rnnt_loss", "joint_logits", "Joint", "get_fused_loss_fn"]

_PREC = {"fp32": _lib.FP32_EXACT, "bf16": _lib.BF16_TC, _lib.FP32_EXACT: _lib.FP32_EXACT, _lib.BF16_TC: _lib.BF16_TC}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _desc(B, T, U, H, V, blank, precision, compact=True, keep=False):
    # torch callers synchronise with the host every step anyway (loss read-back), so the torch surface lets the
    # backward compact ragged batches (one 4-byte read-back); capture-safe callers pass compact=False.
    sync_ok = 1 if (compact and os.environ.get("RNNTB200_COMPACT", "1") != "0"
                    and not torch.cuda.is_current_stream_capturing()) else 0
    return _lib.JointDesc(B, T, U, H, V, int(blank), _PREC[precision],
                          C.c_void_p(torch.cuda.current_stream().cuda_stream).value, sync_ok, 1 if keep else 0)


def _workspace(desc, device):
    sz = C.c_size_t(0)
    _lib.check(_lib.load().rnntb200_joint_workspace_size(C.byref(desc), C.byref(sz)), "rnntb200_joint_workspace_size")
    return torch.empty(sz.value, dtype=torch.uint8, device=device)


def _check(enc, pred, W, b, labels, input_lengths, label_lengths):
    for t, n in ((enc, "enc_acts"), (pred, "pred_acts"), (W, "W"), (b, "b")):
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor: rnnt_b200 has no CPU path" % n)
        if t.dtype != torch.float32:
            raise TypeError("%s must be torch.float32" % n)
    for t, n in ((labels, "labels"), (input_lengths, "input_lengths"), (label_lengths, "label_lengths")):
        if t.dtype != torch.int32:
            raise TypeError("%s must be torch.int32" % n)
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor (device-resident lengths/labels, warprnnt_op.cc:88-94)" % n)
    if enc.dim() != 3 or pred.dim() != 3 or W.dim() != 2 or b.dim() != 1:
        raise ValueError("expected enc (B,T,H), pred (B,U,H), W (H,V), b (V)")
    B, T, H = enc.shape
    if pred.shape[0] != B or pred.shape[2] != H or W.shape[0] != H or b.shape[0] != W.shape[1]:
        raise ValueError("inconsistent shapes")
    U = pred.shape[1]
    if labels.dim() != 2 or labels.shape[0] != B or (labels.shape[1] != U - 1 and U > 1):
        raise ValueError("labels must be (B, U-1)")
    if input_lengths.shape != (B,) or label_lengths.shape != (B,):
        raise ValueError("must have a length per example.")


class _JointRNNT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, pred, W, b, labels, input_lengths, label_lengths, blank, precision, compact, keep=None):
        L = _lib.load()
        enc, pred, W, b = (t.contiguous() for t in (enc, pred, W, b))
        labels, input_lengths, label_lengths = (t.contiguous() for t in (labels, input_lengths, label_lengths))
        _check(enc, pred, W, b, labels, input_lengths, label_lengths)
        B, T, H = enc.shape
        U, V = pred.shape[1], W.shape[1]
        lab = labels if labels.numel() else torch.zeros(1, dtype=torch.int32, device=enc.device)
        with torch.cuda.device(enc.device):
            # a backward will follow: let the forward keep its softmax numerators / tanh outputs in the workspace
            keep = any(ctx.needs_input_grad[:4]) if keep is None else bool(keep)
            desc = _desc(B, T, U, H, V, blank, precision, compact, keep)
            ws = _workspace(desc, enc.device)
            costs = torch.empty(B, dtype=torch.float32, device=enc.device)
            st = L.rnntb200_joint_loss_forward(C.byref(desc), _ptr(enc), _ptr(pred), _ptr(W), _ptr(b), _ptr(lab),
                                               _ptr(label_lengths), _ptr(input_lengths), _ptr(costs), _ptr(ws))
        _lib.check(st, "rnntb200_joint_loss_forward")
        ctx.save_for_backward(enc, pred, W, b, lab, input_lengths, label_lengths)
        ctx.ws, ctx.dims, ctx.blank, ctx.precision, ctx.compact, ctx.keep = ws, (B, T, U, H, V), blank, precision, compact, keep
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        L = _lib.load()
        enc, pred, W, b, lab, input_lengths, label_lengths = ctx.saved_tensors
        B, T, U, H, V = ctx.dims
        g = grad_costs.to(torch.float32).contiguous()
        d_enc, d_pred, dW, db = (torch.empty_like(t) for t in (enc, pred, W, b))
        with torch.cuda.device(enc.device):
            desc = _desc(B, T, U, H, V, ctx.blank, ctx.precision, ctx.compact, ctx.keep)
            st = L.rnntb200_joint_loss_backward(C.byref(desc), _ptr(enc), _ptr(pred), _ptr(W), _ptr(b), _ptr(lab),
                                                _ptr(label_lengths), _ptr(input_lengths), _ptr(g), _ptr(d_enc),
                                                _ptr(d_pred), _ptr(dW), _ptr(db), _ptr(ctx.ws))
        _lib.check(st, "rnntb200_joint_loss_backward")
        ctx.ws = None
        return d_enc, d_pred, dW, db, None, None, None, None, None, None, None


def joint_rnnt_loss(enc_acts, pred_acts, W, b, labels, input_lengths, label_lengths, blank=0, precision="bf16",
                    compact=True, keep_activations=None):
    """Per-utterance RNN-T NLL (B,) of logits = tanh(enc_acts[:,:,None]+pred_acts[:,None]) @ W + b,
    differentiable w.r.t. enc_acts, pred_acts, W, b -- without ever materialising (B,T,U,V).
    precision: 'bf16' (tcgen05 tensor cores, fp32 accumulate) or 'fp32' (exact CUDA-core path).
    compact: let the bf16 backward skip padding tiles of ragged batches (one 4-byte host read-back per chunk);
    pass False for a fully sync-free call (stream capture of the whole step is not validated yet, tools/graph_capture.py).
    keep_activations: None = automatically when a gradient is required (the bf16 forward then leaves its softmax
    numerators (fp16) and tanh outputs in the workspace and the backward is one streaming pass + two GEMMs);
    False = the backward recomputes the projection on the tensor cores (2 bytes per logit less workspace traffic,
    one more tensor-core pass)."""
    return _JointRNNT.apply(enc_acts, pred_acts, W, b, labels, input_lengths, label_lengths, blank, precision, compact,
                            keep_activations)


def joint_logits(enc_acts, pred_acts, W, b):
    """model.py:158-166's literal output: logits (B,T,U,V) float32, materialised (not differentiable;
    use joint_rnnt_loss for training).  For callers that want the reference's tensor itself."""
    L = _lib.load()
    enc, pred, W, b = (t.contiguous() for t in (enc_acts, pred_acts, W, b))
    B, T, H = enc.shape
    U, V = pred.shape[1], W.shape[1]
    with torch.cuda.device(enc.device):
        desc = _desc(B, T, U, H, V, 0, "fp32")
        ws = _workspace(desc, enc.device)
        out = torch.empty(B, T, U, V, dtype=torch.float32, device=enc.device)
        st = L.rnntb200_joint_logits(C.byref(desc), _ptr(enc), _ptr(pred), _ptr(W), _ptr(b), _ptr(out), _ptr(ws))
    _lib.check(st, "rnntb200_joint_logits")
    return out


class Joint(torch.nn.Module):
    """The joint network of model.py:158-166 with Keras-layout parameters:
    dense_1 kernel (P,H) + bias (H) with tanh, dense_2 kernel (H,V) + bias (V).

    ``forward(inp_enc, pred_outputs)`` returns the reference's logits tensor (drop-in for the Keras
    model output); ``loss(...)`` runs the fused path where that tensor never exists."""

    def __init__(self, proj_size, joint_net_size, vocab_size, precision="bf16", blank=0):
        super().__init__()
        self.kernel_1 = torch.nn.Parameter(torch.empty(proj_size, joint_net_size))
        self.bias_1 = torch.nn.Parameter(torch.zeros(joint_net_size))
        self.kernel_2 = torch.nn.Parameter(torch.empty(joint_net_size, vocab_size))
        self.bias_2 = torch.nn.Parameter(torch.zeros(vocab_size))
        torch.nn.init.xavier_uniform_(self.kernel_1)   # Keras Dense default glorot_uniform
        torch.nn.init.xavier_uniform_(self.kernel_2)
        self.precision, self.blank = precision, blank

    def hoist(self, inp_enc, pred_outputs):
        """Dense-1 is linear before its tanh, so it is applied to the (B,T,P) and (B,U,P) inputs
        instead of the (B,T,U,P) lattice (two small library GEMMs): SURVEY 8a2."""
        return inp_enc @ self.kernel_1 + self.bias_1, pred_outputs @ self.kernel_1

    def forward(self, inp_enc, pred_outputs):
        enc_acts, pred_acts = self.hoist(inp_enc, pred_outputs)
        return joint_logits(enc_acts, pred_acts, self.kernel_2, self.bias_2)

    def step(self, f, g):
        """Greedy-decode joint, utils/decoding.py:6-18: ``joint(model, f, g)`` adds ``f`` (B,T,P) to the LAST
        prediction-network frame ``g[:, -1, :]`` and returns ``outputs[:, 0, 0, :]`` -- the (B,V) logits of frame 0."""
        enc_acts, pred_acts = self.hoist(f[:, :1, :], g[:, -1:, :])
        return joint_logits(enc_acts, pred_acts, self.kernel_2, self.bias_2)[:, 0, 0, :]

    def loss(self, inp_enc, pred_outputs, labels, input_lengths, label_lengths):
        enc_acts, pred_acts = self.hoist(inp_enc, pred_outputs)
        return joint_rnnt_loss(enc_acts, pred_acts, self.kernel_2, self.bias_2, labels, input_lengths, label_lengths,
                               self.blank, self.precision)


def get_fused_loss_fn(reduction_factor, joint):
    """utils/loss.get_loss_fn-shaped adapter for the fused path: the returned function takes the joint
    network's INPUTS in place of y_pred (run_rnnt.py:269-273 collapsed into one call)."""
    _lib.load()

    def _loss_fn(y_true, inp_enc, pred_outputs, spec_lengths, label_lengths):
        y_true = y_true.to(torch.int32).contiguous()
        enc_lengths = encoder_lengths(spec_lengths, reduction_factor).contiguous()
        return joint.loss(inp_enc, pred_outputs, y_true, enc_lengths, label_lengths.to(torch.int32).contiguous())

    return _loss_fn
