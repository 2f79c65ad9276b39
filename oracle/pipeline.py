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


# ---- end-to-end triangle-id accounting (BASELINE.json north_star: "bit-exact, apart from a reported count") ----
def classify_mismatches(v_ref, v_got, triangles, tid_ref, tid_got, near_tie, res, edge_eps=1e-3):
  """Why do two triangle-id buffers differ?  v_ref / v_got: the float32 raster vertices [3N] of the CPU chain and of
  the device chain (they differ in the last ulp for some vertices); tid_*: winning triangle per pixel (-1 = none).
  A mismatching pixel is EXPLAINED when it is a depth near-tie (two nearest depths within 1 ulp, `near_tie`), or an
  EDGE FLIP: it lies within `edge_eps` (barycentric units, float64) of an edge of one of the two winners and a corner
  of that winner differs between the two vertex sets -- the inside test of mesh_core.cpp:23-50 flipped because a
  vertex moved by an ulp.  Returns counts; `unexplained` must be 0."""
  tid_ref = np.asarray(tid_ref).reshape(-1)
  tid_got = np.asarray(tid_got).reshape(-1)
  near_tie = np.asarray(near_tie).reshape(-1).astype(bool)
  tri = np.asarray(triangles).reshape(-1, 3)
  vr = np.asarray(v_ref, dtype=np.float32).reshape(-1, 3)
  vg = np.asarray(v_got, dtype=np.float32).reshape(-1, 3)
  moved = np.any(vr.view(np.uint32) != vg.view(np.uint32), axis=1)
  out = {'tri_id_mismatch_px': 0, 'near_tie_px': 0, 'edge_flip_px': 0, 'unexplained_px': 0}
  for p in np.nonzero(tid_ref != tid_got)[0]:
    out['tri_id_mismatch_px'] += 1
    if near_tie[p]:
      out['near_tie_px'] += 1
      continue
    x, y = float(p % res), float(p // res)
    explained = False
    for t in (int(tid_ref[p]), int(tid_got[p])):
      if t < 0:
        continue
      a, b, c = tri[t]
      if not (moved[a] or moved[b] or moved[c]):
        continue
      for vv in (vr, vg):
        p0, p1, p2 = vv[a, :2].astype(np.float64), vv[b, :2].astype(np.float64), vv[c, :2].astype(np.float64)
        e0, e1, e2 = p2 - p0, p1 - p0, np.array([x, y]) - p0
        d00, d01, d11, d02, d12 = e0 @ e0, e0 @ e1, e1 @ e1, e0 @ e2, e1 @ e2
        den = d00 * d11 - d01 * d01
        if den == 0:
          explained = True
          continue
        u, v = (d11 * d02 - d01 * d12) / den, (d00 * d12 - d01 * d02) / den
        if min(abs(u), abs(v), abs(1.0 - u - v)) < edge_eps:
          explained = True
    out['edge_flip_px' if explained else 'unexplained_px'] += 1
  return out
