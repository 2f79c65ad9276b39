"""Host logic of vp_model_create: the vertex tiles (voicepuppet_b200/csrc/topology.cu) must encode
exactly the adjacency Compute_norm walks (reference utils/reconstruct_mesh.py:35-52).  The tables
are pulled through the C ABI (no GPU needed) and the vertex kernel's normal computation is
re-enacted from them in numpy, then compared with the oracle."""
import ctypes

import numpy as np
import pytest

from oracle import reconstruct_oracle as orc
from voicepuppet_b200 import _lib, synthetic

TILE_V, TILE_LV, TILE_LT = 128, 256, 512


def build(model):
  lib = _lib.lib()
  nver = model.meanshape.size // 3
  tri = np.ascontiguousarray((model.tri - 1).astype(np.int32))
  pb = np.ascontiguousarray((model.point_buf - 1).astype(np.int32))
  xyz = np.ascontiguousarray(model.meanshape.reshape(-1, 3).astype(np.float64))
  h = ctypes.c_void_p()
  _lib.check(lib.vp_topology_build(ctypes.byref(h), nver, tri.shape[0], _lib.ptr(tri), _lib.ptr(pb), _lib.ptr(xyz)))
  nt, nl, nh = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
  _lib.check(lib.vp_topology_sizes(h, ctypes.byref(nt), ctypes.byref(nl), ctypes.byref(nh)))
  t = dict(v_int2orig=np.zeros(nver, np.int32), tri_int=np.zeros((tri.shape[0], 4), np.int32),
           tiles=np.zeros((nt.value, 7), np.int32), ltri=np.zeros(nl.value, np.uint32),
           halo=np.zeros(nh.value, np.int32), ring=np.zeros((nver, 8), np.uint16),
           fan=np.zeros((nver, 5), np.uint32))
  _lib.check(lib.vp_topology_copy(h, *[_lib.ptr(t[k]) for k in ('v_int2orig', 'tri_int', 'tiles', 'ltri', 'halo', 'ring', 'fan')]))
  t['slot_off'] = np.zeros(nt.value, np.int32)
  t['slot_tab'] = np.zeros(max(lib.vp_topology_slot_count(h), 1), np.uint16)
  t['fan_slot'] = np.zeros((nver, 5), np.uint32)
  _lib.check(lib.vp_topology_copy_slots(h, _lib.ptr(t['slot_off']), _lib.ptr(t['slot_tab']), _lib.ptr(t['fan_slot'])))
  t['own_tri_off'] = np.zeros(nt.value + 1, np.int32)
  t['own_ltri'] = np.zeros(max(tri.shape[0], 1), np.uint32)
  t['tri_by_orig'] = np.zeros((max(tri.shape[0], 1), 4), np.int32)
  t['fused_ok'] = lib.vp_topology_copy_owned(h, _lib.ptr(t['own_tri_off']), _lib.ptr(t['own_ltri']), _lib.ptr(t['tri_by_orig']))
  lib.vp_topology_destroy(h)
  return t, tri, pb


def fan_normals(fan, pos, nv):
  """What vertex_fan_kernel does: sum over the set mask bits of (u_i - v) x (u_i+1 - v)."""
  fan = fan.astype(np.int64)
  off = np.stack([fan[:, 0] & 0xFFFF, fan[:, 0] >> 16, fan[:, 1] & 0xFFFF, fan[:, 1] >> 16, fan[:, 2] & 0xFFFF,
                  fan[:, 2] >> 16, fan[:, 3] & 0xFFFF, fan[:, 3] >> 16, fan[:, 4] & 0xFFFF], axis=1)
  assert np.all(off % 16 == 0)
  u = off // 16
  mask = fan[:, 4] >> 16
  acc = np.zeros((nv, 3))
  v = pos[:nv]
  for i in range(8):
    on = ((mask >> i) & 1).astype(bool)
    acc[on] += np.cross(pos[u[on, i]] - v[on], pos[u[on, i + 1]] - v[on])
  return acc


def normals_from_tiles(t, shape_orig, force_generic=False):
  """What the vertex kernels do for the normals (fan tiles: vertex_fan_kernel, the others:
  vertex_tile_kernel), in float64."""
  nver = shape_orig.shape[0]
  shape_int = shape_orig[t['v_int2orig']]
  out = np.zeros((nver, 3))
  for v_begin, nv, nlv, nlt, halo_off, ltri_off, is_fan in t['tiles']:
    local = np.concatenate([np.arange(v_begin, v_begin + nv), t['halo'][halo_off:halo_off + nlv - nv]])
    pos = shape_int[local]
    if is_fan and not force_generic:
      acc = fan_normals(t['fan'][v_begin:v_begin + nv], pos, nv)
      with np.errstate(invalid='ignore', divide='ignore'):
        out[t['v_int2orig'][v_begin:v_begin + nv]] = acc / np.linalg.norm(acc, axis=1, keepdims=True)
      continue
    lt = t['ltri'][ltri_off:ltri_off + nlt]
    a, b, c = lt & 1023, (lt >> 10) & 1023, (lt >> 20) & 1023
    fn = np.cross(pos[a] - pos[b], pos[b] - pos[c]) if nlt else np.zeros((0, 3))
    fn = np.concatenate([fn, np.zeros((1, 3))])
    ring = t['ring'][v_begin:v_begin + nv].astype(np.int64)
    ring[ring == 0xFFFF] = nlt
    acc = np.zeros((nv, 3))
    for s in range(8):
      acc = acc + fn[ring[:, s]]
    with np.errstate(invalid='ignore', divide='ignore'):
      out[t['v_int2orig'][v_begin:v_begin + nv]] = acc / np.linalg.norm(acc, axis=1, keepdims=True)
  return out


def check_model(model):
  t, tri, pb = build(model)
  nver, ntri = model.meanshape.size // 3, tri.shape[0]
  # permutations
  assert sorted(t['v_int2orig'].tolist()) == list(range(nver))
  assert sorted(t['tri_int'][:, 3].tolist()) == list(range(ntri))
  o2i = np.empty(nver, np.int64)
  o2i[t['v_int2orig']] = np.arange(nver)
  assert np.array_equal(t['tri_int'][:, :3], o2i[tri[t['tri_int'][:, 3]]])     # corner order preserved
  # tiles partition the vertices and respect the kernel's shared-memory limits
  tiles = t['tiles']
  assert tiles[0, 0] == 0 and np.array_equal(tiles[1:, 0], np.cumsum(tiles[:-1, 1])) and tiles[:, 1].sum() == nver
  assert tiles[:, 1].max() <= TILE_V and tiles[:, 2].max() <= TILE_LV and tiles[:, 3].max() <= TILE_LT
  assert np.all(tiles[:, 2] >= tiles[:, 1])
  # the normals computed from the tables equal Compute_norm
  coeff = synthetic.make_coeffs(1, seed=5)
  shape = orc.shape_formation(coeff[:, :80], coeff[:, 80:144], model)
  want = orc.compute_norm(shape, model)[0]
  for force_generic in (False, True):   # the ring tables stay valid for fan tiles too
    got = normals_from_tiles(t, shape[0].astype(np.float64), force_generic)
    both_nan = np.isnan(want) & np.isnan(got)
    assert np.array_equal(np.isnan(want), np.isnan(got))
    assert np.allclose(np.where(both_nan, 0, got), np.where(both_nan, 0, want), rtol=0, atol=1e-12)
  return t


def test_small_model(small_model):
  check_model(small_model)


def test_full_model_tiles(full_model):
  t = check_model(full_model)
  tiles = t['tiles']
  # spatial order works: tiles are nearly full and the halo stays small
  assert tiles.shape[0] <= 300
  assert tiles[:, 2].mean() < 220
  assert tiles[:, 6].all()        # a manifold mesh: every tile takes the fan path


def test_awkward_meshes():
  """Isolated vertices, valence-8 fans, pad slots in the middle of a ring, duplicate ring entries."""
  rng = np.random.Generator(np.random.PCG64(3))
  nver, ntri = 700, 1500
  tri = rng.integers(0, nver - 20, (ntri, 3))            # last 20 vertices are isolated
  pb = np.full((nver, 8), ntri, dtype=np.int64)
  fill = np.zeros(nver, dtype=np.int64)
  for f in range(ntri):
    for v in tri[f]:
      if fill[v] < 8:
        pb[v, fill[v]] = f
        fill[v] += 1
  pb[5, 0], pb[5, 3] = ntri, pb[5, 0]                     # pad slot first
  pb[6, 1] = pb[6, 0]                                     # duplicate face
  pts = rng.random((nver, 3))
  model = synthetic.SyntheticBFM(
      meanshape=pts.reshape(1, -1).astype(np.float64), idBase=np.zeros((3 * nver, 80), np.float32),
      exBase=np.zeros((3 * nver, 64), np.float32), meantex=np.zeros((1, 3 * nver), np.float32),
      texBase=np.zeros((3 * nver, 80), np.float32), point_buf=(pb + 1).astype(np.float64),
      tri=(tri + 1).astype(np.float64), keypoints=np.arange(68, dtype=np.int32))
  t = check_model(model)
  assert not t['tiles'][:, 6].all()  # a triangle soup does not chain into fans: generic tiles


def test_rejects_bad_indices():
  lib = _lib.lib()
  tri = np.array([[0, 1, 7]], np.int32)
  pb = np.zeros((3, 8), np.int32)
  xyz = np.zeros((3, 3))
  h = ctypes.c_void_p()
  rc = lib.vp_topology_build(ctypes.byref(h), 3, 1, _lib.ptr(tri), _lib.ptr(pb), _lib.ptr(xyz))
  assert rc != 0 and b'out of range' in lib.vp_last_error()


def test_slot_tables_keep_the_normals_and_halve_the_bank_conflicts(full_model):
  """The optional slot tables (VPB200_VERTEX_SLOTS=1): positions staged at slot_tab[local] and gathered through
  fan_slot give the same normals, and in the quarter-warp bank model (tools/bank_conflict_sim.py, which matches
  ncu's conflict count for the default tables) the excess wavefronts drop by more than a third."""
  import os, sys
  sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
  from bank_conflict_sim import excess_wavefronts
  t, tri, pb = build(full_model)
  coeff = synthetic.make_coeffs(1, seed=5)
  shape = orc.shape_formation(coeff[:, :80], coeff[:, 80:144], full_model)[0].astype(np.float64)
  want = orc.compute_norm(shape[None], full_model)[0]
  shape_int = shape[t['v_int2orig']]
  got = np.zeros_like(want)
  for ti, (v_begin, nv, nlv, nlt, halo_off, ltri_off, is_fan) in enumerate(t['tiles']):
    assert is_fan and t['slot_off'][ti] >= 0
    slots = t['slot_tab'][t['slot_off'][ti]:t['slot_off'][ti] + nlv].astype(np.int64)
    assert len(set(slots.tolist())) == nlv and slots.max() < nlv + 8            # a valid placement in s_pos
    local = np.concatenate([np.arange(v_begin, v_begin + nv), t['halo'][halo_off:halo_off + nlv - nv]])
    staged = np.zeros((nlv + 8, 3))
    staged[slots] = shape_int[local]                                            # what the kernel's stage() writes
    # fan_normals reads pos[u] - pos[:nv] with v at local index: own vertex v sits at slot slots[v]
    fan = t['fan_slot'][v_begin:v_begin + nv].astype(np.int64)
    off = np.stack([fan[:, 0] & 0xFFFF, fan[:, 0] >> 16, fan[:, 1] & 0xFFFF, fan[:, 1] >> 16, fan[:, 2] & 0xFFFF,
                    fan[:, 2] >> 16, fan[:, 3] & 0xFFFF, fan[:, 3] >> 16, fan[:, 4] & 0xFFFF], axis=1) // 16
    mask = fan[:, 4] >> 16
    v = staged[slots[:nv]]
    acc = np.zeros((nv, 3))
    for i in range(8):
      on = ((mask >> i) & 1).astype(bool)
      acc[on] += np.cross(staged[off[on, i]] - v[on], staged[off[on, i + 1]] - v[on])
    got[t['v_int2orig'][v_begin:v_begin + nv]] = acc / np.linalg.norm(acc, axis=1, keepdims=True)
  assert np.allclose(got, want, rtol=0, atol=1e-12)
  before, _, _ = excess_wavefronts(t['tiles'][::9], t['fan'])
  after, _, _ = excess_wavefronts(t['tiles'][::9], t['fan_slot'])
  assert after < 0.65 * before, (before, after)


def test_triangle_ownership_tables_of_the_fused_kernel():
  """csrc/fused.cu: every triangle is owned by exactly one tile (the one holding its smallest internal vertex),
  its corners decode -- through the owner's local numbering -- to the triangle's own internal vertices, and the
  resolve pass's table maps an ORIGINAL triangle index to the same corners."""
  for model in (synthetic.make_model(420, 48), synthetic.cached_model()):
    t, tri, pb = build(model)
    assert t['fused_ok'] == 1
    ntri = tri.shape[0]
    off = t['own_tri_off']
    assert off[0] == 0 and off[-1] == ntri and np.all(np.diff(off) >= 0) and np.diff(off).max() <= TILE_LT
    orig2int = np.empty(model.meanshape.size // 3, np.int64)
    orig2int[t['v_int2orig']] = np.arange(orig2int.size)
    for ti, (v_begin, nv, nlv, nlt, halo_off, ltri_off, is_fan) in enumerate(t['tiles']):
      local = np.concatenate([np.arange(v_begin, v_begin + nv), t['halo'][halo_off:halo_off + nlv - nv]])
      rows = np.arange(off[ti], off[ti + 1])
      tri_int = t['tri_int'][rows]
      smallest = tri_int[:, :3].min(axis=1)
      assert np.all((smallest >= v_begin) & (smallest < v_begin + nv))
      lt = t['own_ltri'][rows]
      corners = np.stack([lt & 1023, (lt >> 10) & 1023, (lt >> 20) & 1023], axis=1)
      assert corners.max(initial=0) < nlv
      assert np.array_equal(local[corners], tri_int[:, :3])
    assert np.array_equal(np.sort(t['tri_int'][:, 3]), np.arange(ntri))          # every original triangle once
    assert np.array_equal(t['tri_by_orig'][:, :3], orig2int[tri])


def test_inconsistent_point_buf_disables_the_fused_kernel():
  """A point_buf row that omits one of the vertex's faces: the owner tile may not have staged all the corners
  of a triangle it owns, so the fused kernel must be refused (the separate kernels take the mesh)."""
  model = synthetic.make_model(420, 48)
  t, tri, pb = build(model)
  assert t['fused_ok'] == 1
  import copy
  broken = copy.copy(model)
  broken.point_buf = np.array(model.point_buf, copy=True)
  broken.point_buf[:, :] = tri.shape[0] + 1        # every row padded out: no vertex lists any face
  t2, _, _ = build(broken)
  assert t2['fused_ok'] == 0
