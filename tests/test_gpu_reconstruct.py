"""GPU parity, reconstruction: the drop-in reconstruct_mesh (CUDA, through the C ABI) against the
numpy oracle (pinned to the live reference) and the golden outputs of the reference itself.
Tolerance (BASELINE.json north_star): 1e-5 relative to each array's scale.
Reference: utils/reconstruct_mesh.py:5-223."""
import numpy as np
import pytest

from oracle import reconstruct_oracle as orc
from voicepuppet_b200 import reconstruct_mesh as rm, synthetic

pytestmark = pytest.mark.gpu

REL = 1e-5
NAMES7 = ('shape', 'texture', 'color', 'projection', 'zbuffer', 'landmarks', 'translation')


def rel_err(a, b):
  a = np.asarray(a, dtype=np.float64)
  b = np.asarray(b, dtype=np.float64)
  assert a.shape == b.shape, (a.shape, b.shape)
  return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def test_small_model_against_reference_golden(golden_small, small_model):
  g = golden_small
  for t in range(4):
    c = g['coeffs'][t:t + 1]
    out = rm.Reconstruction(c, small_model)
    assert len(out) == 7
    for name, arr in zip(NAMES7, out):
      ref = g['rec%d_%s' % (t, name)]
      assert arr.shape == ref.shape and arr.dtype == ref.dtype, name
      assert rel_err(arr, ref) <= REL, (name, rel_err(arr, ref))
    out = rm.Reconstruction_rotation(c, small_model, g['jitter'][t])
    assert len(out) == 6
    for name, arr in zip(NAMES7[:6], out):
      ref = g['rot%d_%s' % (t, name)]
      assert arr.shape == ref.shape and arr.dtype == ref.dtype, name
      assert rel_err(arr, ref) <= REL, (name, rel_err(arr, ref))


def test_stage_functions_against_reference_golden(golden_small, small_model):
  g = golden_small
  c = g['coeffs'][0:1]
  idc, exc, texc, ang, gam, trans = rm.Split_coeff(c)
  sh = rm.Shape_formation(idc, exc, small_model)
  want = orc.shape_formation(idc, exc, small_model)
  assert sh.dtype == want.dtype and rel_err(sh, want) <= REL
  tex = rm.Texture_formation(texc, small_model)
  want_tex = orc.texture_formation(texc, small_model)
  assert tex.dtype == want_tex.dtype and rel_err(tex, want_tex) <= REL
  nrm = rm.Compute_norm(want, small_model)
  assert nrm.dtype == np.float64 and rel_err(nrm, g['stage_norm']) <= REL
  rot = rm.Compute_rotation_matrix(ang)
  assert np.array_equal(rot, g['stage_rotation'])
  pr, zb = rm.Projection_layer(want, rot, trans)
  assert rel_err(pr, g['stage_projection']) <= REL and rel_err(zb, g['stage_zbuffer']) <= REL
  col, lit = rm.Illumination_layer(want_tex, nrm, gam)
  wcol, wlit = orc.illumination_layer(want_tex, g['stage_norm'], gam)
  assert rel_err(col, wcol) <= REL and rel_err(lit, wlit) <= REL


@pytest.mark.parametrize('ex_dtype', [np.float64, np.float32])
def test_full_model_against_oracle(full_model, ex_dtype):
  model = full_model if ex_dtype == np.float64 else synthetic.cached_model(ex_dtype=np.float32)
  coeffs = synthetic.make_coeffs(30, seed=1)
  jit = orc.jitter_angle_sequence(30)
  for t in (0, 29):
    got = rm.Reconstruction_rotation(coeffs[t:t + 1], model, jit[t])
    want = orc.reconstruction_rotation(coeffs[t:t + 1], model, jit[t])
    for name, a, b in zip(NAMES7[:6], got, want):
      assert a.dtype == b.dtype, name
      assert rel_err(a, b) <= REL, (name, rel_err(a, b))
    # the projected vertices are much better than the contract: a few 1e-6 pixels
    assert np.max(np.abs(got[3] - want[3])) < 2e-5
  got = rm.Reconstruction(coeffs[3:4], model)
  want = orc.reconstruction(coeffs[3:4], model)
  for name, a, b in zip(NAMES7, got, want):
    assert a.dtype == b.dtype and rel_err(a, b) <= REL, name


def test_full_model_against_reference_golden(golden_full, full_model):
  g = golden_full
  coeffs = synthetic.make_coeffs(30, seed=1)
  jit = orc.jitter_angle_sequence(30)
  for t, res in zip(g['frames'], g['resolutions']):
    out = rm.Reconstruction_rotation(coeffs[t:t + 1], full_model, jit[t])
    key = 'f%d_r%d_' % (t, res)
    for name, arr in zip(NAMES7[:6], out):
      sub = arr if name == 'landmarks' else arr[:, ::int(g['stride'])]
      assert rel_err(sub, g[key + name]) <= REL, name


def test_identity_rotation_shape(small_model, golden_small):
  # SURVEY 8c identity 3: Reconstruction_rotation(...)[0] == Shape_formation(...) @ R(angles)
  c = golden_small['coeffs'][1:2]
  a = golden_small['jitter'][1]
  shape = rm.Reconstruction_rotation(c, small_model, a)[0]
  want = np.matmul(rm.Shape_formation(c[:, :80], c[:, 80:144], small_model).astype(np.float64),
                   rm.Compute_rotation_matrix(a))
  assert rel_err(shape, want) <= 1e-12


def test_linearity_of_the_expression_basis(full_model):
  """Size-independent property: shape(ex1 + ex2) - shape(0) == (shape(ex1) - shape(0)) + (shape(ex2) - shape(0))."""
  c = synthetic.make_coeffs(3, seed=4)
  idc = c[0:1, :80]
  z = np.zeros((1, 64), np.float32)
  s0 = rm.Shape_formation(idc, z, full_model).astype(np.float64)
  s1 = rm.Shape_formation(idc, c[1:2, 80:144], full_model) - s0
  s2 = rm.Shape_formation(idc, c[2:3, 80:144], full_model) - s0
  s12 = rm.Shape_formation(idc, c[1:2, 80:144] + c[2:3, 80:144], full_model) - s0
  assert np.max(np.abs(s12 - (s1 + s2))) < 2e-7


@pytest.mark.parametrize('t', [16, 75, 130])
def test_tensor_core_basis_matches_fp32_and_fp64(full_model, t):
  """K1: the tcgen05 3xTF32 GEMM against the FP32 SIMT kernel and a float64 numpy contraction
  (Shape_formation's expression einsum, reconstruct_mesh.py:21-22)."""
  from voicepuppet_b200 import _lib
  from voicepuppet_b200.model import DeviceModel
  dm = DeviceModel.of(full_model)
  coeffs = synthetic.make_coeffs(t, seed=3)
  dm.set_identity(coeffs[0:1, :80], coeffs[0:1, 144:224])
  base = dm.get_base_shape()
  eye = np.tile(np.eye(3).reshape(1, 9), (t, 1))
  z3, z27 = np.zeros((t, 3), np.float32), np.zeros((t, 27), np.float32)
  out = {}
  try:
    for mode in (1, 2):
      _lib.check(_lib.lib().vp_set_basis_mode(dm.handle, mode))
      out[mode] = dm.reconstruct(coeffs[:, 80:144], eye, z3, z27, want=('shape',))['shape']
  finally:
    _lib.check(_lib.lib().vp_set_basis_mode(dm.handle, 0))
  want = base[None] + np.einsum('ij,tj->ti', full_model.exBase.astype(np.float32).astype(np.float64),
                                coeffs[:, 80:144].astype(np.float64)).reshape(t, -1, 3)
  disp_scale = float(np.abs(want - base[None]).max())
  for mode in (1, 2):
    err = float(np.abs(out[mode] - want).max())
    assert err < 4e-7 * max(1.0, disp_scale), (mode, err)      # fp32-grade accuracy of the displacement
  assert float(np.abs(out[1] - out[2]).max()) < 4e-7


@pytest.mark.parametrize('t', [129, 300, 1100])
def test_tensor_core_basis_many_frame_blocks(full_model, t):
  """K1 above 128 frames: ONE launch walks (frame block, row tile) items and re-splits the frame coefficients when a
  CTA enters the next block.  Checked through the C ABI's device-pointer entry (vp_basis_dev) against a float64
  contraction of the same float32 basis, and against the FP32 kernel; rows of the last, partial block included."""
  import torch
  from voicepuppet_b200 import _lib
  from voicepuppet_b200.model import DeviceModel
  dm = DeviceModel.of(full_model)
  lib = _lib.lib()
  rows_pad = lib.vp_model_rows_pad(dm.handle)
  dev = torch.device('cuda', 0)
  ex = synthetic.make_coeffs(t, seed=11)[:, 80:144].copy()
  ex_d = torch.from_numpy(ex).to(dev)
  st = torch.cuda.current_stream(dev).cuda_stream
  out = {}
  try:
    for mode in (1, 2):
      _lib.check(lib.vp_set_basis_mode(dm.handle, mode))
      disp = torch.full((t, rows_pad), float('nan'), device=dev)
      _lib.check(lib.vp_basis_dev(dm.handle, ex_d.data_ptr(), disp.data_ptr(), t, st))
      torch.cuda.synchronize()
      out[mode] = disp.cpu().numpy()
  finally:
    _lib.check(lib.vp_set_basis_mode(dm.handle, 0))
  nrow = int(full_model.meanshape.size)        # 3 N rows; the rest of rows_pad is padding
  assert np.isfinite(out[2][:, :nrow]).all()
  # same multiset of values as the float64 contraction (the device rows are Morton-renumbered: compare sorted rows)
  want = np.einsum('ij,tj->ti', full_model.exBase.astype(np.float32).astype(np.float64), ex.astype(np.float64))
  scale = float(np.abs(want).max())
  assert float(np.abs(out[1][:, :nrow] - out[2][:, :nrow]).max()) < 4e-7 * max(1.0, scale)
  for k in (0, 127, 128, t // 2, t - 1):
    got = np.sort(out[2][k, :nrow].astype(np.float64))
    assert float(np.abs(got - np.sort(want[k])).max()) < 4e-7 * max(1.0, scale), k


def test_generic_vertex_kernel_matches_fan_kernel(full_model):
  """K2 has two flavours (fan records / ring of staged face normals); the full model is manifold, so it
  normally takes the fan path: force the generic kernel and compare both with the oracle."""
  from voicepuppet_b200 import _lib
  from voicepuppet_b200.model import DeviceModel
  dm = DeviceModel.of(full_model)
  lib = _lib.lib()
  assert lib.vp_model_fan_tiles(dm.handle) == lib.vp_model_ntiles(dm.handle)
  coeffs = synthetic.make_coeffs(2, seed=9)
  jit = orc.jitter_angle_sequence(2)
  want = orc.reconstruction_rotation(coeffs[1:2], full_model, jit[1])
  fan = rm.Reconstruction_rotation(coeffs[1:2], full_model, jit[1])
  try:
    _lib.check(lib.vp_set_vertex_mode(dm.handle, 1))
    gen = rm.Reconstruction_rotation(coeffs[1:2], full_model, jit[1])
  finally:
    _lib.check(lib.vp_set_vertex_mode(dm.handle, 0))
  for name, a, b, c in zip(NAMES7[:6], fan, gen, want):
    assert rel_err(a, c) <= REL and rel_err(b, c) <= REL, name
  assert rel_err(fan[2], gen[2]) <= 2e-6      # colours: same normals up to float32 summation order


def test_awkward_mesh_takes_the_generic_path():
  """A triangle soup with isolated vertices, duplicate and misplaced point_buf entries does not chain into
  fans: the generic kernel must reproduce Compute_norm's slot-order ring sum (NaN normals included)."""
  from voicepuppet_b200 import _lib
  from voicepuppet_b200.model import DeviceModel
  rng = np.random.Generator(np.random.PCG64(3))
  nver, ntri = 700, 1500
  tri = rng.integers(0, nver - 20, (ntri, 3))
  pb = np.full((nver, 8), ntri, dtype=np.int64)
  fill = np.zeros(nver, dtype=np.int64)
  for f in range(ntri):
    for v in tri[f]:
      if fill[v] < 8:
        pb[v, fill[v]] = f
        fill[v] += 1
  pb[5, 0], pb[5, 3] = ntri, pb[5, 0]
  pb[6, 1] = pb[6, 0]
  pb[7, 2] = pb[400, 0]                                   # a face that does not contain the vertex
  pts = rng.random((nver, 3)) * 0.5
  model = synthetic.SyntheticBFM(
      meanshape=pts.reshape(1, -1).astype(np.float32), idBase=(rng.standard_normal((3 * nver, 80)) * 1e-3).astype(np.float32),
      exBase=(rng.standard_normal((3 * nver, 64)) * 1e-3).astype(np.float32),
      meantex=np.full((1, 3 * nver), 128, np.float32), texBase=rng.standard_normal((3 * nver, 80)).astype(np.float32),
      point_buf=(pb + 1).astype(np.float64), tri=(tri + 1).astype(np.float64), keypoints=np.arange(68, dtype=np.int32))
  dm = DeviceModel.of(model)
  lib = _lib.lib()
  assert lib.vp_model_fan_tiles(dm.handle) < lib.vp_model_ntiles(dm.handle)
  c = synthetic.make_coeffs(1, seed=2)
  got = rm.Reconstruction(c, model)
  want = orc.reconstruction(c, model)
  for name, a, b in zip(NAMES7, got, want):
    assert np.array_equal(np.isnan(a), np.isnan(b)), name
    assert rel_err(np.nan_to_num(a), np.nan_to_num(b)) <= REL, (name, rel_err(np.nan_to_num(a), np.nan_to_num(b)))
  assert np.isnan(want[2]).any()                          # isolated vertices: 0/0 normals -> NaN colours


def test_slot_placement_of_the_fan_kernel_is_invisible(full_model):
  """The fan kernel stages local vertex i at a bank-conflict-aware shared-memory slot (default) or at slot i
  (vp_set_vertex_mode(m, 2)): same arithmetic, only the placement differs, so the frames must be byte-identical."""
  from voicepuppet_b200 import _lib, render
  from voicepuppet_b200.model import DeviceModel
  dm = DeviceModel.of(full_model)
  coeffs = synthetic.make_coeffs(20, seed=1)
  want = np.asarray(render.render_sequence(coeffs, full_model, res=224)).copy()
  _lib.check(_lib.lib().vp_set_vertex_mode(dm.handle, 2))
  try:
    got = np.asarray(render.render_sequence(coeffs, full_model, res=224)).copy()
  finally:
    _lib.check(_lib.lib().vp_set_vertex_mode(dm.handle, 0))
  assert want.any() and np.array_equal(got, want)


