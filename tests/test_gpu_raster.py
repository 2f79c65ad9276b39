"""GPU parity, rasterizer: the drop-in mesh_core_cython (CUDA, through the C ABI) against the oracle
and the golden outputs of the reference's Cython module -- bit for bit.
Reference: utils/cython/mesh_core.cpp:108-231."""
import numpy as np
import pytest

from oracle.raster import Oracle, fresh_color_buffers
from voicepuppet_b200 import mesh_core_cython as mc

pytestmark = pytest.mark.gpu


def bits(a):
  return np.ascontiguousarray(a).view(np.uint32)


def gpu_colors(verts, tris, cols, h, w, init=None, want_tid=False):
  c = cols.shape[1] if cols.ndim == 2 else 3
  image, mask, depth = fresh_color_buffers(h, w, c)
  if init is not None and init.size:
    depth = init.astype(np.float32).copy()
  v = np.ascontiguousarray(verts.reshape(-1), dtype=np.float32)
  t = np.ascontiguousarray(tris.reshape(-1), dtype=np.int32)
  col = np.ascontiguousarray(cols.reshape(-1), dtype=np.float32)
  if want_tid:
    tid = mc.render_colors_with_triangle_id(image, mask, v, t, col, depth, t.size // 3, h, w, c)
    return image, mask, depth, tid
  assert mc.render_colors_core(image, mask, v, t, col, depth, t.size // 3, h, w, c) is None
  return image, mask, depth


def cpu_colors(verts, tris, cols, h, w, init=None):
  c = cols.shape[1] if cols.ndim == 2 else 3
  image, mask, depth = fresh_color_buffers(h, w, c)
  if init is not None and init.size:
    depth = init.astype(np.float32).copy()
  tid = np.zeros(h * w, dtype=np.int32)
  t = np.ascontiguousarray(tris.reshape(-1), dtype=np.int32)
  Oracle.render_colors(image, mask, np.ascontiguousarray(verts.reshape(-1), dtype=np.float32), t,
                       np.ascontiguousarray(cols.reshape(-1), dtype=np.float32), depth, t.size // 3, h, w, c,
                       triangle_out=tid)
  return image, mask, depth, tid


def gpu_tri(verts, tris, h, w, init=None):
  depth = np.full((h, w), -99999.0, dtype=np.float32)
  if init is not None and init.size:
    depth = init.astype(np.float32).reshape(h, w).copy()
  tbuf = np.full((h, w), -1, dtype=np.int32)
  wgt = np.zeros((h, 3 * w), dtype=np.float32)
  v2 = np.ascontiguousarray(verts.reshape(-1, 3), dtype=np.float32)
  t2 = np.ascontiguousarray(tris.reshape(-1, 3), dtype=np.int32)
  assert mc.rasterize_triangles_core(v2, t2, depth, tbuf, wgt, v2.shape[0], t2.shape[0], h, w) is None
  return depth, tbuf, wgt


def cpu_tri(verts, tris, h, w, init=None):
  depth = np.full((h, w), -99999.0, dtype=np.float32)
  if init is not None and init.size:
    depth = init.astype(np.float32).reshape(h, w).copy()
  tbuf = np.full((h, w), -1, dtype=np.int32)
  wgt = np.zeros((h, 3 * w), dtype=np.float32)
  v2 = np.ascontiguousarray(verts.reshape(-1, 3), dtype=np.float32)
  t2 = np.ascontiguousarray(tris.reshape(-1, 3), dtype=np.int32)
  Oracle.rasterize_triangles(v2, t2, depth, tbuf, wgt, v2.shape[0], t2.shape[0], h, w)
  return depth, tbuf, wgt


@pytest.mark.parametrize('case', ['lattice', 'special', 'big_c1'])
def test_edge_cases_against_reference_golden(golden_edges, case):
  g = golden_edges
  h, w = int(g['h']), int(g['w'])
  verts, tris, cols, init = (g[case + s] for s in ('_vertices', '_triangles', '_colors', '_init_depth'))
  image, mask, depth = gpu_colors(verts, tris, cols, h, w, init)
  assert np.array_equal(image, g[case + '_image'])
  assert np.array_equal(mask, g[case + '_mask'])
  assert np.array_equal(bits(depth), bits(g[case + '_depth']))
  d2, t2, w2 = gpu_tri(verts, tris, h, w, init)
  assert np.array_equal(t2, g[case + '_tri_id'])
  assert np.array_equal(bits(d2), bits(g[case + '_tri_depth']))
  assert np.array_equal(bits(w2), bits(g[case + '_tri_weight']))


def test_full_frames_against_reference_golden(golden_full, full_model):
  g = golden_full
  tris = (full_model.tri - 1).astype(np.int32)
  for t, res in zip(g['frames'], g['resolutions']):
    key = 'f%d_r%d_' % (t, res)
    verts = g[key + 'vertices']
    cols = g[key + 'colors'].astype(np.float32).reshape(-1, 3)
    image, mask, depth, tid = gpu_colors(verts, tris, cols, res, res, want_tid=True)
    assert np.array_equal(image, g[key + 'image'])
    assert np.array_equal(mask, g[key + 'mask'])
    assert np.array_equal(bits(depth), bits(g[key + 'depth']))
    want_tid = cpu_colors(verts, tris, cols, res, res)[3]
    assert np.array_equal(tid, want_tid)                      # implied triangle id, bit exact
  d2, t2, w2 = gpu_tri(g['f0_r224_vertices'], tris, 224, 224)
  assert np.array_equal(t2, g['f0_tri_id'])
  assert np.array_equal(bits(d2), bits(g['f0_tri_depth']))
  assert np.array_equal(bits(w2), bits(g['f0_tri_weight']))


@pytest.mark.parametrize('res', [512, 1024])
def test_full_frame_high_resolution_matches_oracle(golden_full, full_model, res):
  tris = (full_model.tri - 1).astype(np.int32)
  verts = golden_full['f0_r224_vertices'].reshape(-1, 3).copy()
  verts[:, :2] *= np.float32(res / 224.0)
  cols = golden_full['f0_r224_colors'].astype(np.float32).reshape(-1, 3)
  a = gpu_colors(verts, tris, cols, res, res, want_tid=True)
  b = cpu_colors(verts, tris, cols, res, res)
  for x, y in zip(a, b):
    assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
  assert a[1].mean() > 60                                       # the face covers a good part of the frame


@pytest.mark.parametrize('seed', range(8))
def test_random_soups_match_oracle(seed):
  rng = np.random.Generator(np.random.PCG64(200 + seed))
  h, w = int(rng.integers(8, 90)), int(rng.integers(8, 90))
  nt = int(rng.integers(1, 600))
  nv = 3 * nt
  extent = [1.5, 5.0, 40.0, 200.0][seed % 4]
  centre = rng.random((nt, 1, 3)) * np.array([w + 8, h + 8, 4]) - np.array([4, 4, 0])
  verts = (centre + (rng.random((nt, 3, 3)) - 0.5) * np.array([extent, extent, 1.0])).reshape(nv, 3).astype(np.float32)
  if seed % 2:
    verts[:, :2] = np.round(verts[:, :2] * 2) / 2               # exact ties and on-edge pixel centres
    verts[:, 2] = np.round(verts[:, 2])
  tris = np.arange(nv, dtype=np.int32).reshape(nt, 3)
  share = rng.random(nt) < 0.3
  tris[share] = rng.integers(0, nv, (int(share.sum()), 3))
  c = [3, 1, 4][seed % 3]
  cols = rng.integers(0, 256, (nv, c)).astype(np.float32)
  a = gpu_colors(verts, tris, cols, h, w, want_tid=True)
  b = cpu_colors(verts, tris, cols, h, w)
  for x, y in zip(a, b):
    assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
  a = gpu_tri(verts, tris, h, w)
  b = cpu_tri(verts, tris, h, w)
  for x, y in zip(a, b):
    assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_empty_and_degenerate_inputs():
  verts = np.zeros((3, 3), dtype=np.float32)
  cols = np.zeros((3, 3), dtype=np.float32)
  image, mask, depth = gpu_colors(verts, np.zeros((0, 3), np.int32), cols, 8, 8)
  assert not image.any() and not mask.any() and np.all(depth == np.float32(-99999.0))
  # zero-area triangle on a pixel centre: inverDeno = 0 makes u = v = 0, which the reference's test
  # (u >= 0 && v >= 0 && u + v < 1, mesh_core.cpp:49) accepts -- one pixel is drawn
  verts = np.array([[2, 2, 1], [2, 2, 1], [2, 2, 1]], dtype=np.float32)
  tri = np.array([[0, 1, 2]], np.int32)
  a = gpu_colors(verts, tri, cols, 8, 8, want_tid=True)
  b = cpu_colors(verts, tri, cols, 8, 8)
  for x, y in zip(a, b):
    assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
  assert a[1].sum() == 255 and a[1].reshape(8, 8)[2, 2] == 255


def test_idempotent_second_pass(golden_full, full_model):
  """Rendering the same mesh again into the produced depth buffer changes nothing (strict '>')."""
  tris = np.ascontiguousarray((full_model.tri - 1).astype(np.int32).reshape(-1))
  verts = golden_full['f29_r256_vertices']
  cols = golden_full['f29_r256_colors'].astype(np.float32)
  image, mask, depth = fresh_color_buffers(256, 256, 3)
  mc.render_colors_core(image, mask, verts, tris, cols, depth, tris.size // 3, 256, 256, 3)
  i2, m2, d2 = image.copy(), mask.copy(), depth.copy()
  i2[:] = 7
  mc.render_colors_core(i2, m2, verts, tris, cols, d2, tris.size // 3, 256, 256, 3)
  assert np.all(i2 == 7) and np.array_equal(m2, mask) and np.array_equal(bits(d2), bits(depth))
