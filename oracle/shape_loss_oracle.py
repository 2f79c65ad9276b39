"""TEST INFRASTRUCTURE -- numpy restatement of BFMNet's vertex loss, reference
voicepuppet/bfmnet/bfmnet.py:215-268 (Shape_formation :215-227, add_cost_function :240-268, without the network's
regularisation term :269).  Only tests/ may import this.

PARITY UNPINNED: the reference evaluates these lines inside a TensorFlow-1 graph and TensorFlow is not available
here, so this restatement follows the source line by line but could not be run against the live
reference; it is evaluated in float64 and the GPU op is compared with it under a tolerance.
"""
import numpy as np


def shape_formation(bfm_coeffs, model):
  """bfmnet.py:215-227: [A,144+] coefficients -> [1, A*N, 3] (float64 here)."""
  id_base = np.asarray(model.idBase, dtype=np.float64)
  ex_base = np.asarray(model.exBase, dtype=np.float64)
  mean = np.asarray(model.meanshape, dtype=np.float64)
  c = np.asarray(bfm_coeffs, dtype=np.float64)
  face_shape = np.einsum('ij,aj->ai', id_base, c[:, :80]) + np.einsum('ij,aj->ai', ex_base, c[:, 80:144]) + mean
  face_shape = face_shape.reshape(1, -1, 3)
  return face_shape - np.mean(mean.reshape(1, -1, 3), axis=1, keepdims=True)


def cost(output_ex, bfm_coeffs, seq_len, model, mouth_mask):
  """bfmnet.py:240-267.  output_ex [B,T,64], bfm_coeffs [B,T,257], seq_len [B], mouth_mask [N,3]."""
  b, t = output_ex.shape[0], output_ex.shape[1]
  n3 = np.asarray(model.meanshape).size
  out_coeffs = np.concatenate([bfm_coeffs[:, :, :80], output_ex], axis=-1).reshape(-1, 144)
  output_face_shape = shape_formation(out_coeffs, model).reshape(b, -1, n3)
  face_shape = shape_formation(bfm_coeffs.reshape(-1, bfm_coeffs.shape[-1]), model).reshape(b, -1, n3)
  tmax = int(np.max(seq_len))
  assert tmax == t, 'the reference pads every batch to max(seq_len)'
  vertice_mask = np.tile(mouth_mask.reshape(1, 1, n3), (b, tmax, 1)).astype(np.float64)
  coeff_mask = (np.arange(tmax)[None, :] < np.asarray(seq_len)[:, None]).astype(np.float64)
  diff = np.sum(np.abs(face_shape - output_face_shape) * vertice_mask, axis=-1)
  loss = np.mean(np.sum(diff * coeff_mask, axis=-1))
  video_mask = (np.arange(tmax - 1)[None, :] < (np.asarray(seq_len) - 1)[:, None]).astype(np.float64)
  video_diff = (output_face_shape[:, 1:, :] - output_face_shape[:, :-1, :]) - (face_shape[:, 1:, :] - face_shape[:, :-1, :])
  video_diff = np.sum(np.abs(video_diff) * vertice_mask[:, :-1, :], axis=-1)
  return loss + np.mean(np.sum(video_diff * video_mask, axis=-1))
