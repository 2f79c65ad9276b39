"""The product's rasterizer arithmetic header (voicepuppet_b200/csrc/vp_math.cuh -- what the CUDA kernels include)
compiled for the host (tests/hostcheck/hostcheck.cpp, g++ -ffp-contract=off) and checked bit-for-bit against the
reference golden vectors and the oracle, without a GPU: bounding boxes with the x86 cast emulation, the float32
inside test, flat / interpolated depth, 64-bit z-buffer keys (order, ties, NaN / inf / pre-filled depth), flat colours."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.raster import Oracle
from test_oracle_raster import bits, run_colors, run_tri

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostcheck', 'hostcheck.cpp')
HDR = os.path.join(os.path.dirname(HERE), 'voicepuppet_b200', 'csrc', 'vp_math.cuh')
LIB = os.path.join(HERE, 'hostcheck', '_build', 'libhostcheck.so')

_f32p, _i32p, _u8p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_ubyte)


def _p(a, ty):
  return a.ctypes.data_as(ty)


@pytest.fixture(scope='module')
def host():
  os.makedirs(os.path.dirname(LIB), exist_ok=True)
  if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
    subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fno-fast-math', '-x', 'c++', '-fPIC', '-shared',
                    SRC, '-o', LIB], check=True)
  lib = ctypes.CDLL(LIB)
  lib.hc_clip_trunc_byte.argtypes = [ctypes.c_float]
  lib.hc_clip_trunc_byte.restype = ctypes.c_uint
  lib.hc_depth_code.argtypes = [ctypes.c_float]
  lib.hc_depth_code.restype = ctypes.c_uint

  class Host(object):
    @staticmethod
    def render_colors(image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c, triangle_out=None):
      lib.hc_render_colors(_p(image, _u8p), _p(face_mask, _u8p), _p(vertices, _f32p), _p(triangles, _i32p),
                           _p(colors, _f32p), _p(depth_buffer, _f32p),
                           None if triangle_out is None else _p(triangle_out, _i32p), ntri, h, w, c)

    @staticmethod
    def rasterize_triangles(vertices, triangles, depth_buffer, triangle_buffer, barycentric_weight, nver, ntri, h, w):
      lib.hc_rasterize_triangles(_p(vertices, _f32p), _p(triangles, _i32p), _p(depth_buffer, _f32p),
                                 _p(triangle_buffer, _i32p), _p(barycentric_weight, _f32p), ntri, h, w)

  Host.lib = lib
  return Host


@pytest.mark.parametrize('case', ['lattice', 'special', 'big_c1'])
def test_edge_cases_match_reference_golden(host, golden_edges, case):
  g = golden_edges
  h, w = int(g['h']), int(g['w'])
  verts, tris, cols, init = (g[case + s] for s in ('_vertices', '_triangles', '_colors', '_init_depth'))
  image, mask, depth = run_colors(host, verts, tris, cols, h, w, init)
  assert np.array_equal(image, g[case + '_image']) and np.array_equal(mask, g[case + '_mask'])
  assert np.array_equal(bits(depth), bits(g[case + '_depth']))
  d2, t2, w2 = run_tri(host, verts, tris, h, w, init)
  assert np.array_equal(t2, g[case + '_tri_id'])
  assert np.array_equal(bits(d2), bits(g[case + '_tri_depth'])) and np.array_equal(bits(w2), bits(g[case + '_tri_weight']))


@pytest.mark.parametrize('seed', range(6))
def test_random_soups_match_oracle(host, seed):
  rng = np.random.Generator(np.random.PCG64(900 + seed))
  h, w = int(rng.integers(8, 70)), int(rng.integers(8, 70))
  nt = int(rng.integers(1, 400))
  nv = 3 * nt
  extent = [1.5, 5.0, 40.0][seed % 3]
  centre = rng.random((nt, 1, 3)) * np.array([w + 8, h + 8, 4]) - np.array([4, 4, 0])
  verts = (centre + (rng.random((nt, 3, 3)) - 0.5) * np.array([extent, extent, 1.0])).reshape(nv, 3).astype(np.float32)
  if seed % 2:
    verts[:, :2] = np.round(verts[:, :2] * 2) / 2
    verts[:, 2] = np.round(verts[:, 2])
  if seed == 4:
    verts[::17, 0] = np.nan
    verts[5::23, 1] = np.inf
    verts[7::29, 2] = np.nan
    verts[11::31, 0] = 3e9
  tris = np.arange(nv, dtype=np.int32).reshape(nt, 3)
  cols = rng.integers(0, 256, (nv, 3)).astype(np.float32)
  for a, b in zip(run_colors(Oracle, verts, tris, cols, h, w), run_colors(host, verts, tris, cols, h, w)):
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
  for a, b in zip(run_tri(Oracle, verts, tris, h, w), run_tri(host, verts, tris, h, w)):
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_full_frame_matches_reference_golden(host, golden_full, full_model):
  g = golden_full
  tris = (full_model.tri - 1).astype(np.int32)
  t, res = int(g['frames'][0]), int(g['resolutions'][0])
  key = 'f%d_r%d_' % (t, res)
  image, mask, depth = run_colors(host, g[key + 'vertices'], tris, g[key + 'colors'].astype(np.float32).reshape(-1, 3), res, res)
  assert np.array_equal(image, g[key + 'image']) and np.array_equal(mask, g[key + 'mask'])
  assert np.array_equal(bits(depth), bits(g[key + 'depth']))


def test_scalar_helpers(host):
  lib = host.lib
  # np.clip(c, 0, 255).astype(int32) of infer_bfmvid.py:98
  for c in (-5.0, -0.0, 0.0, 0.999, 1.0, 127.5, 254.999, 255.0, 300.0, float('nan'), float('inf'), float('-inf')):
    want = 0 if c != c else int(np.clip(c, 0, 255))
    assert lib.hc_clip_trunc_byte(c) == want, c
  # depth codes order like the floats, -0 == +0
  vals = np.array([-np.inf, -99999.0, -1.5, -1e-30, -0.0, 0.0, 1e-30, 2.5, 1e30, np.inf], dtype=np.float32)
  codes = [lib.hc_depth_code(float(v)) for v in vals]
  assert codes[4] == codes[5]
  assert all(a <= b for a, b in zip(codes, codes[1:])) and len(set(codes)) == len(codes) - 1
