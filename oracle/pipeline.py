"""TEST INFRASTRUCTURE -- the reference's per-frame loop on the CPU, end to end.

Follows voicepuppet/pixrefer/infer_bfmvid.py:85-109 (render_face up to the rasterized frame):
Reconstruction_rotation with the jitter angles (or Reconstruction with the coefficient's own
angles), colours clipped and truncated, flat-shaded z-buffer rasterization.  The reconstruction is
oracle/reconstruct_oracle.py (numpy, pinned bit-for-bit to the live reference), the rasterizer is
the reference's own C++ compiled unmodified (oracle/_ref) when present, else the C restatement
(oracle/mesh_core_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import os

import numpy as np

from . import reconstruct_oracle as orc
from .raster import Oracle, Reference, fresh_color_buffers


def rasterizer(prefer_reference=True):
  if prefer_reference and Reference.available():
    return Reference, 'reference'
  return Oracle, 'port'


def frame_raster_inputs(coeff_row, model, angles, res):
  """One frame's reconstruction -> (vertices f32 [3N], colors f32 [3N], recon tuple)."""
  if angles is None:
    out = orc.reconstruction(coeff_row, model)
  else:
    out = orc.reconstruction_rotation(coeff_row, model, np.asarray(angles, dtype=np.float32).reshape(1, 3))
  vertices, colors = orc.raster_inputs(out[3], out[4], out[2], res)
  return vertices, colors, out


def render_frame(coeff_row, model, angles, res, triangles=None, impl=None, want_triangle_id=False):
  """-> (image [res,res,3] u8, mask [res,res] u8, depth [res,res] f32[, triangle_id])."""
  if triangles is None:
    triangles = orc.triangles_flat(model)
  vertices, colors, _ = frame_raster_inputs(coeff_row, model, angles, res)
  image, mask, depth = fresh_color_buffers(res, res, 3)
  ntri = triangles.size // 3
  if want_triangle_id:
    tid = np.zeros(res * res, dtype=np.int32)
    Oracle.render_colors(image, mask, vertices, triangles, colors, depth, ntri, res, res, 3, triangle_out=tid)
    return image.reshape(res, res, 3), mask.reshape(res, res), depth.reshape(res, res), tid.reshape(res, res)
  impl = impl or rasterizer()[0]
  impl.render_colors(image, mask, vertices, triangles, colors, depth, ntri, res, res, 3)
  return image.reshape(res, res, 3), mask.reshape(res, res), depth.reshape(res, res)


def render_sequence(coeffs, model, res=224, angles='jitter', impl=None):
  """CPU twin of voicepuppet_b200.render.render_sequence -> uint8 [T,res,res,3]."""
  coeffs = np.asarray(coeffs, dtype=np.float32)
  t = coeffs.shape[0]
  if isinstance(angles, str):
    angles = orc.jitter_angle_sequence(t)[:, 0, :]
  triangles = orc.triangles_flat(model)
  out = np.zeros((t, res, res, 3), dtype=np.uint8)
  for i in range(t):
    a = None if angles is None else np.asarray(angles)[i]
    out[i] = render_frame(coeffs[i:i + 1], model, a, res, triangles, impl)[0]
  return out


# ---- multi-process CPU baseline (frames are independent) ---------------------------------------
_POOL_STATE = {}


def _pool_init(model_kwargs, res):
  from voicepuppet_b200 import synthetic
  os.environ.setdefault('OMP_NUM_THREADS', '1')
  model = synthetic.cached_model(**model_kwargs)
  _POOL_STATE['model'] = model
  _POOL_STATE['res'] = res
  _POOL_STATE['triangles'] = orc.triangles_flat(model)
  _POOL_STATE['impl'] = rasterizer()[0]


def _pool_work(args):
  coeffs, angles = args
  model, res = _POOL_STATE['model'], _POOL_STATE['res']
  acc = 0
  for i in range(coeffs.shape[0]):
    img = render_frame(coeffs[i:i + 1], model, angles[i], res, _POOL_STATE['triangles'], _POOL_STATE['impl'])[0]
    acc += int(img[::16, ::16].sum())
  return coeffs.shape[0], acc


def timed_pool_run(coeffs, angles, res, workers, model_kwargs=None, repeats=1):
  """Render `coeffs` with `workers` processes; returns (frames, seconds) excluding pool start-up."""
  import multiprocessing as mp
  import time
  model_kwargs = model_kwargs or {}
  ctx = mp.get_context('fork')
  t = coeffs.shape[0]
  bounds = np.linspace(0, t, workers + 1).astype(int)
  jobs = [(coeffs[a:b], angles[a:b]) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
  with ctx.Pool(workers, initializer=_pool_init, initargs=(model_kwargs, res)) as pool:
    pool.map(_pool_work, [(coeffs[:1], angles[:1])] * workers)   # warm every worker (model load, imports)
    best = None
    for _ in range(repeats):
      t0 = time.perf_counter()
      done = pool.map(_pool_work, jobs, chunksize=1)
      dt = time.perf_counter() - t0
      best = dt if best is None else min(best, dt)
  return sum(d[0] for d in done), best
