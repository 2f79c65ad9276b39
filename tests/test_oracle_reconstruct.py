"""Pins oracle/reconstruct_oracle.py against golden outputs of the live reference
(utils/reconstruct_mesh.py:5-223) and, where /root/reference exists, against the reference
module itself bit for bit."""
import numpy as np
import pytest

from oracle import reconstruct_oracle as orc
from voicepuppet_b200 import synthetic

NAMES7 = ('shape', 'texture', 'color', 'projection', 'zbuffer', 'landmarks', 'translation')
# goldens were produced by numpy on another host: float32 einsum/sum order may differ by SIMD width,
# so float32-accumulated quantities get 2e-6 relative to the array scale, float64 ones 1e-12.
def close(a, b, rel):
  scale = max(1.0, float(np.max(np.abs(b))))
  return np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) <= rel * scale


def test_small_model_all_frames(golden_small, small_model):
  g = golden_small
  for t in range(4):
    c = g['coeffs'][t:t + 1]
    out = orc.reconstruction(c, small_model)
    for name, arr in zip(NAMES7, out):
      ref = g['rec%d_%s' % (t, name)]
      assert arr.shape == ref.shape and arr.dtype == ref.dtype, name
      assert close(arr, ref, 2e-6), name
    out = orc.reconstruction_rotation(c, small_model, g['jitter'][t])
    for name, arr in zip(NAMES7[:6], out):
      ref = g['rot%d_%s' % (t, name)]
      assert arr.shape == ref.shape and arr.dtype == ref.dtype, name
      assert close(arr, ref, 2e-6), name


def test_stage_functions(golden_small, small_model):
  g = golden_small
  c = g['coeffs'][0:1]
  sh = orc.shape_formation(c[:, :80], c[:, 80:144], small_model)
  assert close(orc.compute_norm(sh, small_model), g['stage_norm'], 2e-5)
  rot = orc.rotation_matrix(c[:, 224:227])
  assert close(rot, g['stage_rotation'], 1e-7)
  pr, zb = orc.projection_layer(sh, rot, c[:, 254:257])
  assert close(pr, g['stage_projection'], 2e-6) and close(zb, g['stage_zbuffer'], 2e-6)


def test_jitter_sequence_matches_reference_globals(golden_small):
  seq = orc.jitter_angle_sequence(60)
  assert np.array_equal(seq[:4], golden_small['jitter'])
  # triangle wave: +0.005 per frame until |angle| > 0.03, then back (infer_bfmvid.py:85-89)
  assert abs(float(seq[:, 0, 1].max()) - 0.035) < 1e-6 and abs(float(seq[:, 0, 1].min()) + 0.035) < 1e-6
  assert np.all(seq[:, 0, 0] == seq[:, 0, 1]) and np.all(seq[:, 0, 0] == seq[:, 0, 2])


def test_full_model_subsampled(golden_full, full_model):
  g = golden_full
  coeffs = synthetic.make_coeffs(30, seed=1)
  jit = orc.jitter_angle_sequence(30)
  for t, res in zip(g['frames'], g['resolutions']):
    out = orc.reconstruction_rotation(coeffs[t:t + 1], full_model, jit[t])
    key = 'f%d_r%d_' % (t, res)
    for name, arr in zip(NAMES7[:6], out):
      sub = arr if name == 'landmarks' else arr[:, ::int(g['stride'])]
      assert close(sub, g[key + name], 2e-6), name
    verts, cols = orc.raster_inputs(out[3], out[4], out[2], int(res))
    # float32 vertices: identical except where numpy's float32 einsum order differs across hosts
    assert np.mean(verts != g[key + 'vertices']) < 0.2
    assert np.max(np.abs(verts - g[key + 'vertices'])) < 1e-3
    assert np.max(np.abs(cols - g[key + 'colors'].astype(np.float32))) <= 1.0


def test_identity_rotation_shape(small_model, golden_small):
  # SURVEY 8c identity 3: Reconstruction_rotation(...)[0] == Shape_formation(...) @ R(angles)
  c = golden_small['coeffs'][1:2]
  a = golden_small['jitter'][1]
  shape = orc.reconstruction_rotation(c, small_model, a)[0]
  want = np.matmul(orc.shape_formation(c[:, :80], c[:, 80:144], small_model), orc.rotation_matrix(a))
  assert np.array_equal(shape, want)


def test_bit_exact_against_live_reference(live_reference, small_model, golden_small):
  ref = live_reference
  for t in range(4):
    c = golden_small['coeffs'][t:t + 1]
    for x, y in zip(ref.Reconstruction(c, small_model), orc.reconstruction(c, small_model)):
      assert x.dtype == y.dtype and np.array_equal(x, y)
    a = golden_small['jitter'][t]
    for x, y in zip(ref.Reconstruction_rotation(c, small_model, a), orc.reconstruction_rotation(c, small_model, a)):
      assert x.dtype == y.dtype and np.array_equal(x, y)
  # an all-float32 model changes the dtype chain (shape/texture float32, normals float64)
  m32 = synthetic.make_model(300, 40, ex_dtype=np.float32)
  c = synthetic.make_coeffs(1, seed=3)
  for x, y in zip(ref.Reconstruction(c, m32), orc.reconstruction(c, m32)):
    assert x.dtype == y.dtype and np.array_equal(x, y)
