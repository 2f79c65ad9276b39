"""GPU parity for the other two exports of the reference's native module, through the
drop-in voicepuppet_b200.mesh_core_cython: bit-exact against the golden outputs of the reference's Cython
module and against the oracle on seeded random cases.
Reference: utils/cython/mesh_core.cpp:85-105 (get_normal_core), 234-333 (render_texture_core)."""
import numpy as np
import pytest

import texture_cases as tc
from oracle.raster import Oracle
from voicepuppet_b200 import _lib, mesh_core_cython as mc

pytestmark = pytest.mark.gpu


def bits(a):
  return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope='module')
def golden():
  return tc.load()


@pytest.mark.parametrize('name', tc.TEXTURE_CASES)
def test_texture_matches_reference_golden(golden, name):
  case = tc.case_of(golden, name)
  for mapping in (0, 1):
    image, depth = tc.run_texture(mc.render_texture_core, case, mapping)
    assert np.array_equal(bits(image), bits(golden['%s_m%d_image' % (name, mapping)]))
    assert np.array_equal(bits(depth), bits(golden['%s_m%d_depth' % (name, mapping)]))
  image, depth = tc.run_texture(mc.render_texture_core, case, 1, golden[name + '_init_depth'])
  assert np.array_equal(bits(image), bits(golden[name + '_pre_image']))
  assert np.array_equal(bits(depth), bits(golden[name + '_pre_depth']))


@pytest.mark.parametrize('seed', range(8))
def test_random_texture_cases_match_oracle(seed):
  case = tc.random_texture_case(500 + seed)
  for mapping in (0, 1):
    a = tc.run_texture(Oracle.render_texture, case, mapping)
    b = tc.run_texture(mc.render_texture_core, case, mapping)
    for x, y in zip(a, b):
      assert np.array_equal(bits(x), bits(y))


def test_texture_large_frame_matches_oracle(full_model, golden_full):
  """The full mesh at 512x512 with a 256x256 texture: the size the hot path renders at."""
  res = 512
  verts = golden_full['f0_r224_vertices'].reshape(-1, 3).copy()
  verts[:, :2] *= res / 224.0
  tris = np.ascontiguousarray((full_model.tri - 1).astype(np.int32))
  rng = np.random.Generator(np.random.PCG64(77))
  nver = verts.shape[0]
  tex_coords = np.zeros((nver, 3), np.float32)
  tex_coords[:, :2] = rng.random((nver, 2)).astype(np.float32) * 255
  texture = rng.random((256, 256, 3)).astype(np.float32)
  case = dict(vertices=np.ascontiguousarray(verts), triangles=tris, tex_coords=tex_coords, tex_triangles=tris.copy(),
              texture=texture, h=res, w=res, c=3, tex_h=256, tex_w=256, tex_c=3)
  a = tc.run_texture(Oracle.render_texture, case, 1)
  b = tc.run_texture(mc.render_texture_core, case, 1)
  for x, y in zip(a, b):
    assert np.array_equal(bits(x), bits(y))
  assert (b[0] != -1).mean() > 0.3


@pytest.mark.parametrize('name', tc.NORMAL_CASES)
def test_normals_match_reference_golden(golden, name):
  normal = golden[name + '_init'].copy()
  tris = golden[name + '_triangles']
  mc.get_normal_core(normal, golden[name + '_tri_normal'], tris, tris.shape[0])
  assert np.array_equal(bits(normal), bits(golden[name + '_normal']))


@pytest.mark.parametrize('seed', range(6))
def test_random_normal_cases_match_oracle(seed):
  tris, tri_normal, init = tc.random_normal_case(600 + seed)
  a, b = init.copy(), init.copy()
  Oracle.get_normal(a, tri_normal, tris, tris.shape[0])
  mc.get_normal_core(b, tri_normal, tris, tris.shape[0])
  assert np.array_equal(bits(a), bits(b))


def test_normals_full_mesh_and_empty(full_model):
  tris = np.ascontiguousarray((full_model.tri - 1).astype(np.int32))
  rng = np.random.Generator(np.random.PCG64(5))
  tri_normal = rng.standard_normal((tris.shape[0], 3)).astype(np.float32)
  nver = full_model.meanshape.size // 3
  a, b = np.zeros((nver, 3), np.float32), np.zeros((nver, 3), np.float32)
  Oracle.get_normal(a, tri_normal, tris, tris.shape[0])
  mc.get_normal_core(b, tri_normal, tris, tris.shape[0])
  assert np.array_equal(bits(a), bits(b))
  c = np.ones((4, 3), np.float32)
  mc.get_normal_core(c, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), 0)
  assert np.all(c == 1)


def test_bad_indices_are_errors():
  with pytest.raises(_lib.VpError):
    mc.get_normal_core(np.zeros((3, 3), np.float32), np.zeros((1, 3), np.float32), np.array([[0, 1, 3]], np.int32), 1)
  case = tc.random_texture_case(1)
  case['tex_triangles'] = case['tex_triangles'].copy()
  case['tex_triangles'][0, 0] = case['tex_coords'].shape[0]
  with pytest.raises(_lib.VpError):
    tc.run_texture(mc.render_texture_core, case, 0)
