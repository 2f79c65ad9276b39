"""Pins the oracle restatements of render_texture_core / get_normal_core (oracle/mesh_core_oracle.c)
bit-for-bit against golden outputs of the reference's Cython module (tests/golden/make_golden_extra.py)
and against the reference's C++ compiled in place (oracle/_ref), on seeded random cases.
Reference: utils/cython/mesh_core.cpp:85-105, 234-333."""
import numpy as np
import pytest

import texture_cases as tc
from oracle.raster import Oracle, Reference


def bits(a):
  return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope='module')
def golden():
  return tc.load()


@pytest.mark.parametrize('name', tc.TEXTURE_CASES)
@pytest.mark.parametrize('reverse', [False, True])
def test_texture_matches_reference_golden(golden, name, reverse):
  case = tc.case_of(golden, name)
  for mapping in (0, 1):
    image, depth = tc.run_texture(Oracle.render_texture, case, mapping, reverse=reverse)
    assert np.array_equal(bits(image), bits(golden['%s_m%d_image' % (name, mapping)]))
    assert np.array_equal(bits(depth), bits(golden['%s_m%d_depth' % (name, mapping)]))
    assert (image != -1).any()
  image, depth = tc.run_texture(Oracle.render_texture, case, 1, golden[name + '_init_depth'], reverse=reverse)
  assert np.array_equal(bits(image), bits(golden[name + '_pre_image']))
  assert np.array_equal(bits(depth), bits(golden[name + '_pre_depth']))


@pytest.mark.parametrize('name', tc.NORMAL_CASES)
def test_normals_match_reference_golden(golden, name):
  normal = golden[name + '_init'].copy()
  tris = golden[name + '_triangles']
  Oracle.get_normal(normal, golden[name + '_tri_normal'], tris, tris.shape[0])
  assert np.array_equal(bits(normal), bits(golden[name + '_normal']))


@pytest.mark.skipif(not Reference.available(), reason='oracle/_ref not built')
@pytest.mark.parametrize('seed', range(8))
def test_random_texture_cases_match_compiled_reference(seed):
  case = tc.random_texture_case(200 + seed)
  for mapping in (0, 1):
    a = tc.run_texture(Reference.render_texture, case, mapping)
    b = tc.run_texture(Oracle.render_texture, case, mapping, reverse=bool(seed & 1))
    for x, y in zip(a, b):
      assert np.array_equal(bits(x), bits(y))


@pytest.mark.skipif(not Reference.available(), reason='oracle/_ref not built')
@pytest.mark.parametrize('seed', range(6))
def test_random_normal_cases_match_compiled_reference(seed):
  tris, tri_normal, init = tc.random_normal_case(300 + seed)
  a, b = init.copy(), init.copy()
  Reference.get_normal(a, tri_normal, tris, tris.shape[0])
  Oracle.get_normal(b, tri_normal, tris, tris.shape[0])
  assert np.array_equal(bits(a), bits(b))
