"""Drop-in for the evaluation contact sheet of the reference, ``utils/bfm_visual.py:88-154``
(``plot_bfm_coeff_seq``, called by voicepuppet/bfmnet/train_bfmnet.py:130-138 every eval step).

The reference reconstructs and rasterizes up to 30 real + 30 predicted frames one at a time (Reconstruction
+ render_colors_core at 224x224) and tiles them into a 9 x 10 grid; here both sequences go through the batched
GPU path in two calls and only the tiling and the JPEG encoding stay on the host.
"""
import os

import numpy as np

from .render import render_sequence

BLOCK_X, BLOCK_Y, IMG_SIZE = 10, 9, 224   # bfm_visual.py:90-92


def _merge_seq(coeff_seq, facemodel, big_img, time, h_index, render_fn=None):
  """bfm_visual.py:94-130: frames 0..time-1 of the first sequence of the batch, Reconstruction with the
  coefficients' own angles, channel swap (:124), tile (i // 10 + h_index, i % 10)."""
  if time <= 0:
    return big_img
  coeffs = np.ascontiguousarray(np.asarray(coeff_seq)[0, :time, :], dtype=np.float32)
  frames = np.asarray((render_fn or render_sequence)(coeffs, facemodel, res=IMG_SIZE, angles=None))
  for i in range(time):
    r, c = i // BLOCK_X + h_index, i % BLOCK_X
    big_img[r * IMG_SIZE:(r + 1) * IMG_SIZE, c * IMG_SIZE:(c + 1) * IMG_SIZE] = frames[i][:, :, ::-1]
  return big_img


def contact_sheet(facemodel, seq_len, real_bfm_coeff_seq, bfm_coeff_seq, id_coeff=None, texture_coeff=None,
                  render_fn=None):
  """The uint8 [9*224, 10*224, 3] image plot_bfm_coeff_seq writes: rows 0-2 the real sequence, rows 3-5 the
  predicted expression coefficients spliced into the real (or the given) identity / texture.
  ``render_fn(coeffs[T,257], facemodel, res=, angles=)`` replaces the GPU renderer in the CPU tests of the tiling."""
  real = np.asarray(real_bfm_coeff_seq)
  pred = np.asarray(bfm_coeff_seq)
  time = 30 if seq_len[0] > 30 else int(seq_len[0])          # :133-137
  big_img = np.zeros((IMG_SIZE * BLOCK_Y, IMG_SIZE * BLOCK_X, 3), dtype=np.uint8)
  big_img = _merge_seq(real, facemodel, big_img, time, 0, render_fn)
  if id_coeff is None or texture_coeff is None:               # :147-150
    spliced = np.concatenate([real[:, :, :80], pred[:, :, :], real[:, :, 144:]], axis=2)
  else:
    n = real.shape[1]
    spliced = np.concatenate([np.tile(id_coeff, (1, n, 1)), pred[:, :, :], np.tile(texture_coeff, (1, n, 1)),
                              real[:, :, 224:]], axis=2)
  return _merge_seq(spliced, facemodel, big_img, time, 3, render_fn)


def plot_bfm_coeff_seq(save_dir, facemodel, step, seq_len, real_bfm_coeff_seq, bfm_coeff_seq, id_coeff=None,
                       texture_coeff=None):
  """Same signature and output file as bfm_visual.py:88-154: '<save_dir>/bfmnet_<step>.jpg'."""
  big_img = contact_sheet(facemodel, seq_len, real_bfm_coeff_seq, bfm_coeff_seq, id_coeff, texture_coeff)
  path = '{}/bfmnet_{}.jpg'.format(save_dir, step)
  try:
    import cv2
    cv2.imwrite(path, big_img)
  except ImportError:                                         # cv2.imwrite takes BGR: swap for PIL
    from PIL import Image
    Image.fromarray(big_img[:, :, ::-1]).save(path)
  return None
