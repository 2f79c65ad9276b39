"""Synthetic BFM-shaped face model and coefficient sequences.

The real ``BFM_model_front.mat`` (reference ``utils/bfm_load_data.py:9-21``) is not
available offline, so benchmarks and tests use a model with the same attribute
contract: 35709 vertices, 70789 triangles, a disc
topology with 627 boundary vertices, 1-based ``tri`` / ``point_buf`` with the pad
value F+1, spatially smooth bases.

Everything here is built from IEEE-exact operations only (+, -, *, /, sqrt,
integer->float conversion): no sin/cos/exp and no BLAS.  numpy evaluates each
elementwise ufunc call with individually rounded operations, so the arrays are
bit-reproducible on any x86-64 host -- which is what lets golden fixtures made in
one container be checked on another machine.  Normal variates are Irwin-Hall
(sum of 12 uniforms - 6) for the same reason.
"""
import hashlib
import os

import numpy as np

N_VERTICES = 35709
N_BOUNDARY = 627
N_TRIANGLES = 70789

# cos/sin of the golden angle pi*(3-sqrt(5)), as decimal literals (exactly parsed)
_GOLD_C = -0.7373688780783197
_GOLD_S = 0.6754902942615238


class SyntheticBFM(object):
  """Duck-typed stand-in for ``bfm_load_data.BFM`` (same 8 attributes)."""

  def __init__(self, meanshape, idBase, exBase, meantex, texBase, point_buf, tri, keypoints):
    self.meanshape = meanshape
    self.idBase = idBase
    self.exBase = exBase
    self.meantex = meantex
    self.texBase = texBase
    self.point_buf = point_buf
    self.tri = tri
    self.keypoints = keypoints

  def checksum(self):
    h = hashlib.sha256()
    for name in ('meanshape', 'idBase', 'exBase', 'meantex', 'texBase', 'point_buf', 'tri', 'keypoints'):
      a = np.ascontiguousarray(getattr(self, name))
      h.update(name.encode())
      h.update(str(a.dtype).encode())
      h.update(str(a.shape).encode())
      h.update(a.tobytes())
    return h.hexdigest()


def _uniform(rng, shape):
  # (raw >> 11) * 2**-53: exact, identical on every host
  return rng.random(shape)


def normal_ih(rng, shape):
  """Approximately N(0,1): Irwin-Hall(12) - 6, summed in a fixed order."""
  acc = np.zeros(shape, dtype=np.float64)
  for _ in range(12):
    acc = acc + _uniform(rng, shape)
  return acc - 6.0


def _rotation_walk(n, c, s):
  """(cos(k*a), sin(k*a)) for k=1..n by repeated exact-op rotation, renormalised."""
  out = np.empty((n, 2), dtype=np.float64)
  x, y = 1.0, 0.0
  for k in range(n):
    x, y = x * c - y * s, x * s + y * c
    if (k & 63) == 63:
      r = (x * x + y * y) ** 0.5
      x, y = x / r, y / r
    out[k, 0] = x
    out[k, 1] = y
  return out


def _circle_step(n):
  """cos/sin of 2*pi/n from exact ops: half-angle recurrence from cos(pi/2)=0 is not
  general, so use the tangent half-angle of a Newton-refined chord instead.
  For reproducibility we only need *a* fixed step close to 2*pi/n."""
  # rational approximation of 2*pi/n then 12-term Taylor series (exact ops only)
  a = 6.283185307179586 / n
  a2 = a * a
  c = 1.0
  s = a
  tc = 1.0
  ts = a
  for k in range(1, 12):
    tc = -tc * a2 / ((2 * k - 1) * (2 * k))
    ts = -ts * a2 / ((2 * k) * (2 * k + 1))
    c = c + tc
    s = s + ts
  r = (c * c + s * s) ** 0.5
  return c / r, s / r


def _points(n_vertices, n_boundary):
  n_in = n_vertices - n_boundary
  c, s = _circle_step(n_boundary)
  ring = _rotation_walk(n_boundary, c, s)
  spiral = _rotation_walk(n_in, _GOLD_C, _GOLD_S)
  k = np.arange(1, n_in + 1, dtype=np.float64)
  rad = 0.992 * np.sqrt((k - 0.5) / n_in)
  inner = spiral * rad[:, None]
  return np.concatenate([ring, inner], axis=0)


def _cheb_features(x, y, degree):
  """[N, P] products T_a(x) T_b(y), a+b <= degree, by the three-term recurrence."""
  def cheb(u):
    t = [np.ones_like(u), u]
    for _ in range(2, degree + 1):
      t.append(2.0 * u * t[-1] - t[-2])
    return t
  tx, ty = cheb(x), cheb(y)
  cols = []
  for a in range(degree + 1):
    for b in range(degree + 1 - a):
      cols.append(tx[a] * ty[b])
  return cols


def _smooth_basis(cols, rng, n_out, amp):
  """amp * sum_p g[p, j] * cols[p] for j < n_out, accumulated one feature at a time."""
  n = cols[0].shape[0]
  out = np.zeros((n, n_out), dtype=np.float64)
  scale = amp / (len(cols) ** 0.5)
  for col in cols:
    g = normal_ih(rng, (n_out,)) * scale
    out = out + col[:, None] * g[None, :]
  return out


def build_point_buf(tri0, n_vertices, slots=8):
  """Adjacent-face table (0-based in, 1-based out) filled in triangle order; pad = F+1."""
  n_tri = tri0.shape[0]
  pb = np.full((n_vertices, slots), n_tri + 1, dtype=np.float64)
  fill = np.zeros(n_vertices, dtype=np.int64)
  for f in range(n_tri):
    for v in tri0[f]:
      k = fill[v]
      if k >= slots:
        raise ValueError('vertex %d has valence > %d' % (v, slots))
      pb[v, k] = f + 1
      fill[v] = k + 1
  return pb


def make_model(n_vertices=N_VERTICES, n_boundary=N_BOUNDARY, seed=0, ex_dtype=np.float64,
               float_dtype=np.float32, degree=5):
  """Build the synthetic model.  Defaults give V=35709, F=70789 (disc: F = 2V - 2 - B)."""
  from scipy.spatial import Delaunay

  pts = _points(n_vertices, n_boundary)
  x, y = pts[:, 0], pts[:, 1]
  tri0 = Delaunay(pts).simplices.astype(np.int64)
  # orient counter-clockwise in the (x, y) plane
  a, b, c = pts[tri0[:, 0]], pts[tri0[:, 1]], pts[tri0[:, 2]]
  area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
  flip = area2 < 0
  tri0[flip] = tri0[flip][:, [0, 2, 1]]
  # canonical order so that qhull's facet order does not matter
  tri0 = tri0[np.lexsort((tri0[:, 2], tri0[:, 1], tri0[:, 0]))]
  expect = 2 * n_vertices - 2 - n_boundary
  if tri0.shape[0] != expect:
    raise RuntimeError('triangulation gave %d faces, expected %d' % (tri0.shape[0], expect))

  r2 = x * x + y * y
  q = x * x + (y + 0.1) * (y + 0.1)
  bump = 1.0 + q / 0.02
  z = 0.6 * (1.0 - r2) + 0.25 / (bump * bump)
  mean = np.stack([0.8 * x, 1.0 * y, z], axis=1).reshape(1, -1)

  rng = np.random.Generator(np.random.PCG64(seed))
  cols = _cheb_features(x, y, degree)
  n3 = 3 * n_vertices

  def basis(k, amp):
    # [N, 3*k] -> [3N, k] with row = 3*v + axis, as reshape([1,-1,3]) in the reference expects
    raw = _smooth_basis(cols, rng, 3 * k, amp)
    return raw.reshape(n_vertices, 3, k).reshape(n3, k)

  id_base = basis(80, 0.01)
  ex_base = basis(64, 0.01)
  tex_base = basis(80, 8.0)
  meantex = np.full((1, n3), 128.0)

  point_buf = build_point_buf(tri0, n_vertices)
  keypoints = (np.arange(68, dtype=np.int64) * (n_vertices - 1)) // 67

  return SyntheticBFM(
      meanshape=mean.astype(float_dtype),
      idBase=id_base.astype(float_dtype),
      exBase=ex_base.astype(ex_dtype),
      meantex=meantex.astype(float_dtype),
      texBase=tex_base.astype(float_dtype),
      point_buf=point_buf,
      tri=(tri0 + 1).astype(np.float64),
      keypoints=keypoints.astype(np.int32))


_CACHE = {}


def cached_model(n_vertices=N_VERTICES, n_boundary=N_BOUNDARY, seed=0, ex_dtype=np.float64,
                 float_dtype=np.float32, cache_dir=None):
  """make_model with an in-process cache and an optional on-disk .npz cache."""
  key = (n_vertices, n_boundary, seed, np.dtype(ex_dtype).str, np.dtype(float_dtype).str)
  if key in _CACHE:
    return _CACHE[key]
  path = None
  if cache_dir is None:
    cache_dir = os.environ.get('VPB200_CACHE', '/tmp/vpb200_cache')
  if cache_dir:
    path = os.path.join(cache_dir, 'bfm_%d_%d_%d_%s_%s.npz' % (
        n_vertices, n_boundary, seed, np.dtype(ex_dtype).name, np.dtype(float_dtype).name))
  model = None
  if path and os.path.exists(path):
    try:
      with np.load(path) as z:
        model = SyntheticBFM(**{k: z[k] for k in z.files})
    except Exception:
      model = None
  if model is None:
    model = make_model(n_vertices, n_boundary, seed, ex_dtype, float_dtype)
    if path:
      try:
        os.makedirs(cache_dir, exist_ok=True)
        tmp = path + '.%d.tmp.npz' % os.getpid()
        np.savez(tmp, **{k: getattr(model, k) for k in (
            'meanshape', 'idBase', 'exBase', 'meantex', 'texBase', 'point_buf', 'tri', 'keypoints')})
        os.replace(tmp, path)
      except OSError:
        pass
  _CACHE[key] = model
  return model


def make_coeffs(n_frames, seed=1, rho=0.9):
  """[T,257] float32: id/tex/angles/gamma/translation fixed per clip (tiled over frames
  like reference voicepuppet/pixrefer/infer_bfmvid.py:223-224), expression an AR(1)
  sequence with unit marginal variance."""
  rng = np.random.Generator(np.random.PCG64(seed))
  clip = np.zeros(257, dtype=np.float64)
  clip[0:80] = normal_ih(rng, (80,))
  clip[144:224] = normal_ih(rng, (80,))
  clip[224:227] = 0.1 * normal_ih(rng, (3,))
  clip[227:254] = 0.1 * normal_ih(rng, (27,))
  clip[254:257] = 0.05 * normal_ih(rng, (3,))
  out = np.tile(clip[None, :], (n_frames, 1))
  innov = (1.0 - rho * rho) ** 0.5
  ex = normal_ih(rng, (64,))
  for t in range(n_frames):
    if t:
      ex = rho * ex + innov * normal_ih(rng, (64,))
    out[t, 80:144] = ex
  return out.astype(np.float32)
