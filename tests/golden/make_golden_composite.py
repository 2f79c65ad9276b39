"""Generates tests/golden/composite.npz with the calls the reference makes (cv2.cvtColor, cv2.resize, the
numpy paste and the float conversion of voicepuppet/pixrefer/infer_bfmvid.py:111-121,234) on synthetic
rasters.  infer_bfmvid.py itself cannot be imported (TensorFlow), so those lines are driven verbatim here.
Run from the repo root:  python tests/golden/make_golden_composite.py   (needs cv2; OpenCV version stored)."""
import os

import cv2
import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))


def reference_lines(new_image, center_x, center_y, ratio, transform_params, img):
  # infer_bfmvid.py:80-82
  ratio *= transform_params[2]
  tx = -int((transform_params[3] / ratio))
  ty = -int((transform_params[4] / ratio))
  # :111-121
  new_image = cv2.cvtColor(new_image, cv2.COLOR_BGR2RGB)
  new_image = cv2.resize(new_image, (
      int(round(new_image.shape[0] / ratio)), int(round(new_image.shape[1] / ratio))))
  back_new_image = np.zeros((img.shape[0], img.shape[1], img.shape[2]), dtype=img.dtype)
  center_face_x = new_image.shape[1] // 2
  center_face_y = new_image.shape[0] // 2
  ry = center_y - center_face_y + new_image.shape[0] - ty
  rx = center_x - center_face_x + new_image.shape[1] - tx
  back_new_image[center_y - center_face_y - ty:ry, center_x - center_face_x - tx:rx, :] = new_image
  # :234
  face3d = cv2.cvtColor(back_new_image, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
  return back_new_image, face3d


def synthetic_raster(rng, res):
  """A face-like blob: smooth colours inside an ellipse, zeros outside, plus noise."""
  y, x = np.mgrid[0:res, 0:res]
  inside = ((x - res / 2) / (0.42 * res)) ** 2 + ((y - res / 2) / (0.47 * res)) ** 2 < 1
  img = np.stack([128 + 90 * np.sin(x / 17.0 + k) * np.cos(y / 23.0 - k) for k in range(3)], axis=2)
  img += rng.standard_normal(img.shape) * 12
  return (np.clip(img, 0, 255) * inside[:, :, None]).astype(np.uint8)


def main():
  rng = np.random.Generator(np.random.PCG64(41))
  img = np.zeros((512, 512, 3), np.uint8)
  cases = [  # center_x, center_y, ratio, transform_params (w0, h0, s, tx, ty)
      (256, 250, 1.05, [512, 512, 0.97, 12.3, -20.7]),
      (250, 262, 0.80, [512, 512, 1.0, 0.0, 0.0]),       # upscale
      (260, 240, 2.00, [512, 512, 1.0, -30.0, 14.0]),    # exact 2x downscale: OpenCV's area path
      (256, 256, 1.00, [512, 512, 1.0, 5.0, 5.0]),       # no resize
      (200, 300, 1.37, [512, 512, 1.21, 40.5, 33.3]),
      (256, 256, 0.51, [512, 512, 1.0, 0.0, 0.0]),       # large upscale, nearly fills the canvas
  ]
  data = dict(opencv_version=cv2.__version__, n_cases=len(cases))
  for i, (cx, cy, ratio, tp) in enumerate(cases):
    raster = synthetic_raster(rng, 224)
    canvas, face3d = reference_lines(raster, cx, cy, ratio, np.array(tp, dtype=np.float64), img)
    data['c%d_raster' % i] = raster
    data['c%d_args' % i] = np.array([cx, cy, ratio] + tp, dtype=np.float64)
    data['c%d_canvas' % i] = canvas
    data['c%d_face3d_sum' % i] = np.float64(face3d.astype(np.float64).sum())   # the float image is canvas / 255
  path = os.path.join(OUT, 'composite.npz')
  np.savez_compressed(path, **data)
  print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
  main()
