"""Pins oracle/mesh_core_oracle.c (order-independent restatement) bit-for-bit against
  * golden outputs of the reference's Cython module (tests/golden/make_golden.py), and
  * the reference's C++ compiled in place (oracle/_ref), on seeded random soups.
Reference: utils/cython/mesh_core.cpp:108-231."""
import numpy as np
import pytest

from oracle.raster import Oracle, Reference, fresh_color_buffers


def bits(a):
  return np.ascontiguousarray(a).view(np.uint32)


def run_colors(impl, verts, tris, cols, h, w, init=None, **kw):
  c = cols.shape[1] if cols.ndim == 2 else 3
  image, mask, depth = fresh_color_buffers(h, w, c)
  if init is not None and init.size:
    depth = init.astype(np.float32).copy()
  impl.render_colors(image, mask, np.ascontiguousarray(verts.reshape(-1)), np.ascontiguousarray(tris.reshape(-1)),
                     np.ascontiguousarray(cols.reshape(-1)), depth, tris.size // 3, h, w, c, **kw)
  return image, mask, depth


def run_tri(impl, verts, tris, h, w, init=None, **kw):
  depth = np.full((h, w), -99999.0, dtype=np.float32)
  if init is not None and init.size:
    depth = init.astype(np.float32).reshape(h, w).copy()
  tbuf = np.full((h, w), -1, dtype=np.int32)
  wgt = np.zeros((h, 3 * w), dtype=np.float32)
  v2 = np.ascontiguousarray(verts.reshape(-1, 3))
  t2 = np.ascontiguousarray(tris.reshape(-1, 3))
  impl.rasterize_triangles(v2, t2, depth, tbuf, wgt, v2.shape[0], t2.shape[0], h, w, **kw)
  return depth, tbuf, wgt


@pytest.mark.parametrize('case', ['lattice', 'special', 'big_c1'])
@pytest.mark.parametrize('reverse', [False, True])
def test_edge_cases_match_reference_golden(golden_edges, case, reverse):
  g = golden_edges
  h, w = int(g['h']), int(g['w'])
  verts, tris, cols, init = (g[case + s] for s in ('_vertices', '_triangles', '_colors', '_init_depth'))
  image, mask, depth = run_colors(Oracle, verts, tris, cols, h, w, init, reverse=reverse)
  assert np.array_equal(image, g[case + '_image'])
  assert np.array_equal(mask, g[case + '_mask'])
  assert np.array_equal(bits(depth), bits(g[case + '_depth']))
  d2, t2, w2 = run_tri(Oracle, verts, tris, h, w, init, reverse=reverse)
  assert np.array_equal(t2, g[case + '_tri_id'])
  assert np.array_equal(bits(d2), bits(g[case + '_tri_depth']))
  assert np.array_equal(bits(w2), bits(g[case + '_tri_weight']))


def test_full_frames_match_reference_golden(golden_full, full_model):
  g = golden_full
  tris = (full_model.tri - 1).astype(np.int32)
  for t, res in zip(g['frames'], g['resolutions']):
    key = 'f%d_r%d_' % (t, res)
    verts = g[key + 'vertices']
    cols = g[key + 'colors'].astype(np.float32).reshape(-1, 3)
    tid = np.zeros(res * res, dtype=np.int32)
    image, mask, depth = run_colors(Oracle, verts, tris, cols, res, res, triangle_out=tid)
    assert np.array_equal(image, g[key + 'image'])
    assert np.array_equal(mask, g[key + 'mask'])
    assert np.array_equal(bits(depth), bits(g[key + 'depth']))
    assert np.array_equal(tid >= 0, mask > 0)
  verts = g['f0_r224_vertices']
  d2, t2, w2 = run_tri(Oracle, verts, tris, 224, 224)
  assert np.array_equal(t2, g['f0_tri_id'])
  assert np.array_equal(bits(d2), bits(g['f0_tri_depth']))
  assert np.array_equal(bits(w2), bits(g['f0_tri_weight']))


@pytest.mark.skipif(not Reference.available(), reason='oracle/_ref not built')
@pytest.mark.parametrize('seed', range(6))
def test_random_soups_match_compiled_reference(seed):
  rng = np.random.Generator(np.random.PCG64(100 + seed))
  h, w = int(rng.integers(8, 70)), int(rng.integers(8, 70))
  nt = int(rng.integers(1, 400))
  nv = 3 * nt
  extent = [1.5, 5.0, 40.0][seed % 3]                     # sub-pixel, small and frame-sized triangles
  centre = rng.random((nt, 1, 3)) * np.array([w + 8, h + 8, 4]) - np.array([4, 4, 0])
  verts = (centre + (rng.random((nt, 3, 3)) - 0.5) * np.array([extent, extent, 1.0])).reshape(nv, 3)
  verts = verts.astype(np.float32)
  if seed % 2:
    verts[:, :2] = np.round(verts[:, :2] * 2) / 2         # half-integer lattice: ties and on-edge centres
    verts[:, 2] = np.round(verts[:, 2])
  tris = np.arange(nv, dtype=np.int32).reshape(nt, 3)
  share = rng.random(nt) < 0.3                            # some triangles share vertices with others
  tris[share] = rng.integers(0, nv, (int(share.sum()), 3))
  cols = rng.integers(0, 256, (nv, 3)).astype(np.float32)
  a = run_colors(Reference, verts, tris, cols, h, w)
  b = run_colors(Oracle, verts, tris, cols, h, w)
  c = run_colors(Oracle, verts, tris, cols, h, w, reverse=True)
  for x, y, z in zip(a, b, c):
    assert np.array_equal(x.view(np.uint8), y.view(np.uint8)) and np.array_equal(x.view(np.uint8), z.view(np.uint8))
  a = run_tri(Reference, verts, tris, h, w)
  b = run_tri(Oracle, verts, tris, h, w, reverse=bool(seed & 1))
  for x, y in zip(a, b):
    assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_empty_inputs():
  verts = np.zeros((3, 3), dtype=np.float32)
  tris = np.zeros((0, 3), dtype=np.int32)
  cols = np.zeros((3, 3), dtype=np.float32)
  image, mask, depth = run_colors(Oracle, verts, tris, cols, 8, 8)
  assert not image.any() and not mask.any() and np.all(depth == np.float32(-99999.0))


def test_near_tie_report():
  # two coplanar triangles at the same flat depth overlapping one pixel: reported as a tie
  verts = np.array([[0, 0, 1], [6, 0, 1], [0, 6, 1], [6, 6, 1]], dtype=np.float32)
  tris = np.array([[0, 1, 2], [1, 3, 2], [0, 3, 2]], dtype=np.int32)
  ties = Oracle.near_ties(verts.reshape(-1), tris.reshape(-1), 3, 8, 8, ulps=1)
  assert ties.sum() > 0
