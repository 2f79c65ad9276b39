"""Training-time twin of the hot path: BFMNet's vertex loss (reference voicepuppet/bfmnet/bfmnet.py:215-268,
``Shape_formation`` + ``add_cost_function``) as one differentiable op on the GPU.

The reference builds two [B*T, 107127] float32 shapes with tf.einsum (label coefficients; label identity +
predicted expression) and takes mouth-weighted L1 norms of their difference and of its temporal difference.
Identity and mean cancel in that difference, so only the expression contraction of the hot path (K1) is needed:
D = exBase . (ex_label - ex_pred); the backward pass is the transposed contraction.  See csrc/shape_loss.cu.

    loss_fn = ExpressionShapeLoss(facemodel, mouth_mask)        # mouth_mask [35709, 3] float32 (10 on the mouth, 1 elsewhere)
    loss = loss_fn(output_bfm_coeffs, bfm_coeffs[:, :, 80:144], seq_len)   # == loss + video_loss of bfmnet.py:258-267
    loss.backward()

The regularisation term (tf.losses.get_regularization_loss(), :269) belongs to the network, not to this op.
"""
import ctypes

import numpy as np

from . import _lib
from .model import DeviceModel


class ExpressionShapeLoss(object):
  def __init__(self, facemodel, vertex_mask=None, device=0):
    import torch
    self.dm = DeviceModel.of(facemodel, device)
    self.device = torch.device('cuda', device)
    nver = self.dm.nver
    mask = np.ones((nver, 3), np.float32) if vertex_mask is None else np.ascontiguousarray(
        np.asarray(vertex_mask, dtype=np.float32).reshape(nver, 3))
    handle = ctypes.c_void_p()
    _lib.check(_lib.lib().vp_loss_mask_create(self.dm.handle, _lib.ptr(mask), ctypes.byref(handle)))
    self._mask = handle

  def __del__(self):
    try:
      if getattr(self, '_mask', None):
        _lib.lib().vp_loss_mask_destroy(self._mask)
        self._mask = None
    except Exception:
      pass

  def __call__(self, pred_ex, label_ex, seq_len):
    """pred_ex, label_ex: float32 CUDA tensors [B,T,64]; seq_len: [B] ints (list, numpy or tensor).
    Returns a float32 scalar tensor; gradients flow to pred_ex (and to label_ex if it requires them)."""
    import torch
    seq = torch.as_tensor(np.asarray(seq_len.cpu() if hasattr(seq_len, 'cpu') else seq_len), dtype=torch.int32).to(self.device)
    return _ExpressionShapeLossFn.apply(pred_ex, label_ex, seq, self)


def _make_fn():
  import torch

  class Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_ex, label_ex, seq, owner):
      if not (pred_ex.is_cuda and label_ex.is_cuda and pred_ex.dtype == torch.float32 and label_ex.dtype == torch.float32):
        raise ValueError('pred_ex / label_ex must be float32 CUDA tensors')
      if pred_ex.shape != label_ex.shape or pred_ex.dim() != 3 or pred_ex.shape[2] != 64:
        raise ValueError('pred_ex and label_ex must both be [B,T,64]')
      b, t = int(pred_ex.shape[0]), int(pred_ex.shape[1])
      if seq.numel() != b:
        raise ValueError('seq_len must hold one length per sequence')
      delta = (label_ex - pred_ex).contiguous()
      loss = torch.zeros(1, dtype=torch.float64, device=pred_ex.device)
      need_grad = pred_ex.requires_grad or label_ex.requires_grad
      grad = torch.empty_like(delta) if need_grad else None
      stream = ctypes.c_void_p(torch.cuda.current_stream(pred_ex.device).cuda_stream)
      _lib.check(_lib.lib().vp_expression_loss_dev(
          owner.dm.handle, ctypes.c_void_p(delta.data_ptr()), ctypes.c_void_p(seq.data_ptr()), owner._mask, b, t,
          ctypes.c_void_p(loss.data_ptr()), None if grad is None else ctypes.c_void_p(grad.data_ptr()), stream))
      ctx.grad_delta = grad
      return loss.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
      g = ctx.grad_delta
      if g is None:
        return None, None, None, None
      g = g * grad_out
      return (-g if ctx.needs_input_grad[0] else None), (g if ctx.needs_input_grad[1] else None), None, None

  return Fn


class _Lazy(object):
  _fn = None

  def apply(self, *args):
    if _Lazy._fn is None:
      _Lazy._fn = _make_fn()
    return _Lazy._fn.apply(*args)


_ExpressionShapeLossFn = _Lazy()
