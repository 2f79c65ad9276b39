"""TEST INFRASTRUCTURE -- CPU restatement of the post-raster composite of the reference's frame loop
(SURVEY.md section 8f rank 1).  Only tests/ may import this.

What it restates (voicepuppet/pixrefer/infer_bfmvid.py):
  :111     cv2.cvtColor(new_image, COLOR_BGR2RGB)          channel swap
  :112-113 cv2.resize(new_image, (S, S)), S = int(round(224 / ratio))   bilinear, uint8
  :115-121 paste into a zero canvas of the identity image's shape at (center - S // 2 - t)
  :234-236 cv2.cvtColor(face3d, COLOR_BGR2RGB).astype(np.float32) / 255.0 -> inputs[0, ..., 3:6]

cv2.resize is third-party (OpenCV, un-pinned by the reference; 4.13.0 in this image).  Its INTER_LINEAR path
for 8-bit images is fixed point (imgproc/src/resize.cpp: resizeGeneric_ with HResizeLinear / VResizeLinear,
INTER_RESIZE_COEF_BITS = 11), restated in `resize_linear_u8`:
  scale = 1 / (dst / src) in double; f = float((d + 0.5) * scale - 0.5); s = floor(f); f -= s
  x axis: s < 0 -> (s, f) = (0, 0); s >= src - 1 -> (s, f) = (src - 1, 0);  y axis: rows are clipped, f is kept
  coefficients: round-half-even(float32(1 - f) * 2048), round-half-even(f * 2048)
  horizontal: row[dx] = S[sx] * a0 + S[sx + 1] * a1                      (int32)
  vertical:   out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
  exact 2x downscale switches to the 2x2 area mean (a + b + c + d + 2) >> 2 (resize.cpp: INTER_LINEAR with
  iscale == 2 becomes INTER_AREA).
tests/test_oracle_composite.py pins this bit-for-bit against cv2 itself wherever cv2 is importable, and against
tests/golden/composite.npz produced with cv2 in this container.
"""
import numpy as np

COEF_SCALE = 2048


def axis_table(ssize, dsize, is_y):
  """(s0, s1, c0, c1) per destination index, int32."""
  scale = 1.0 / (float(dsize) / float(ssize))
  d = np.arange(dsize, dtype=np.float64)
  f = ((d + 0.5) * scale - 0.5).astype(np.float32)
  s = np.floor(f).astype(np.int64)
  f = (f - s.astype(np.float32)).astype(np.float32)
  if not is_y:
    lo = s < 0
    f[lo] = 0
    s[lo] = 0
    hi = s >= ssize - 1
    f[hi] = 0
    s[hi] = ssize - 1
  c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
  c1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
  s1 = np.clip(s + 1, 0, ssize - 1).astype(np.int32)
  s0 = np.clip(s, 0, ssize - 1).astype(np.int32)
  return s0, s1, c0, c1


def resize_linear_u8(src, dw, dh):
  """cv2.resize(src, (dw, dh)) for uint8 [h, w, c], default interpolation."""
  sh, sw, _ = src.shape
  if dw == sw and dh == sh:
    return src.copy()
  S = src.astype(np.int64)
  if sw == 2 * dw and sh == 2 * dh:
    return ((S[0::2, 0::2] + S[0::2, 1::2] + S[1::2, 0::2] + S[1::2, 1::2] + 2) >> 2).astype(np.uint8)
  sx0, sx1, ax0, ax1 = axis_table(sw, dw, False)
  sy0, sy1, ay0, ay1 = axis_table(sh, dh, True)
  rows = S[:, sx0, :] * ax0[None, :, None] + S[:, sx1, :] * ax1[None, :, None]
  r0, r1 = rows[sy0], rows[sy1]
  out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
  return out.astype(np.uint8)


def placement(center_x, center_y, ratio, transform_params, src=224):
  """infer_bfmvid.py:80-82,112-121: (S, x0, y0) of the pasted face."""
  ratio = ratio * transform_params[2]
  tx = -int((transform_params[3] / ratio))
  ty = -int((transform_params[4] / ratio))
  size = int(round(src / ratio))
  half = size // 2
  return size, center_x - half - tx, center_y - half - ty


def composite(raster, center_x, center_y, ratio, transform_params, canvas_hw):
  """raster uint8 [res, res, 3] (as written by render_colors_core) -> (render_face's return value uint8
  [H, W, 3], the float32 [H, W, 3] that lands in inputs[0, ..., 3:6])."""
  res = raster.shape[0]
  size, x0, y0 = placement(center_x, center_y, ratio, transform_params, res)
  swapped = raster[:, :, ::-1]                                     # :111
  small = resize_linear_u8(np.ascontiguousarray(swapped), size, size)
  canvas = np.zeros((canvas_hw[0], canvas_hw[1], 3), dtype=np.uint8)
  canvas[y0:y0 + size, x0:x0 + size, :] = small                    # :121 (raises like numpy when it does not fit)
  face3d = canvas[:, :, ::-1].astype(np.float32) / 255.0           # :234
  return canvas, face3d
