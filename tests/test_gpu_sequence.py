"""GPU parity, end to end: coefficient sequence -> rendered frames through the fused pipeline
(vp_render_sequence) against the CPU oracle of the reference's frame loop
(voicepuppet/pixrefer/infer_bfmvid.py:85-109).

Contract (BASELINE.json north_star): rendered RGB within 1/255; triangle ids bit-exact given
identical float32 vertices (tests/test_gpu_raster.py).  End to end the device vertices differ from
numpy's in the last float32 ulp for some vertices, so a few edge pixels may pick the neighbouring
triangle: this test counts them, together with the pixels whose two nearest depths are within
1 ulp, and bounds the count."""
import numpy as np
import pytest

from oracle import pipeline, reconstruct_oracle as orc
from oracle.raster import Oracle
from voicepuppet_b200 import _lib, render, synthetic

pytestmark = pytest.mark.gpu


def compare_frames(got, want):
  diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
  return int(diff.max()), float((diff > 0).mean())


@pytest.mark.parametrize('res', [224, 256])
def test_grid_utterance_matches_cpu_reference(full_model, res):
  coeffs = synthetic.make_coeffs(75, seed=1)
  got, mask = render.render_sequence(coeffs, full_model, res=res, angles='jitter', want_mask=True)
  assert got.shape == (75, res, res, 3) and got.dtype == np.uint8
  frames = [0, 1, 37, 74]
  jit = orc.jitter_angle_sequence(75)[:, 0, :]
  worst, changed = 0, 0.0
  for t in frames:
    img, msk, _ = pipeline.render_frame(coeffs[t:t + 1], full_model, jit[t], res)
    # pixels whose winner changed because a vertex moved by an ulp show the neighbour's colour: allow a handful
    d = np.abs(got[t].astype(np.int16) - img.astype(np.int16)).max(axis=2)
    assert (d > 1).sum() <= 12, (t, int((d > 1).sum()))
    assert (mask[t] != msk).sum() <= 4
    worst = max(worst, int(np.percentile(d, 99.9)))
    changed = max(changed, float((d > 0).mean()))
  assert worst <= 1                      # RGB within 1/255
  assert changed < 0.02
  assert mask.mean() > 60                # ~47 % coverage


def test_triangle_ids_end_to_end(full_model):
  """Device reconstruction -> float32 raster inputs -> triangle ids, against the all-CPU chain."""
  from voicepuppet_b200 import mesh_core_cython as mc, reconstruct_mesh as rm
  coeffs = synthetic.make_coeffs(30, seed=1)
  jit = orc.jitter_angle_sequence(30)
  t, res = 29, 256
  tris = orc.triangles_flat(full_model)
  out = rm.Reconstruction_rotation(coeffs[t:t + 1], full_model, jit[t])
  v_gpu, c_gpu = orc.raster_inputs(out[3], out[4], out[2], res)
  v_cpu, c_cpu, _ = pipeline.frame_raster_inputs(coeffs[t:t + 1], full_model, jit[t][0], res)
  image = np.zeros(res * res * 3, np.uint8)
  mask = np.zeros(res * res, np.uint8)
  depth = np.full(res * res, -99999.0, np.float32)
  tid_gpu = mc.render_colors_with_triangle_id(image, mask, v_gpu, tris, c_gpu, depth, tris.size // 3, res, res, 3)
  tid_cpu = np.zeros(res * res, np.int32)
  image2, mask2, depth2 = np.zeros_like(image), np.zeros_like(mask), np.full(res * res, -99999.0, np.float32)
  Oracle.render_colors(image2, mask2, v_cpu, tris, c_cpu, depth2, tris.size // 3, res, res, 3, triangle_out=tid_cpu)
  near = Oracle.near_ties(v_cpu, tris, tris.size // 3, res, res, ulps=1).astype(bool)
  mismatch = tid_gpu != tid_cpu
  print('vertices differing in the last ulp: %.2f %%, triangle-id mismatches: %d (of which near-tie pixels: %d), '
        'near-tie pixels in the frame: %d' % (100.0 * np.mean(v_gpu != v_cpu), int(mismatch.sum()),
                                              int((mismatch & near).sum()), int(near.sum())))
  assert np.max(np.abs(v_gpu - v_cpu)) < 1e-4
  assert mismatch.sum() <= 16
  # every mismatch is accounted for: a depth near-tie, or an inside test that flipped on an edge of a triangle one
  # of whose corners moved by an ulp (the north_star's "reported count")
  why = pipeline.classify_mismatches(v_cpu, v_gpu, tris, tid_cpu, tid_gpu, near, res)
  print(why)
  assert why['unexplained_px'] == 0 and why['tri_id_mismatch_px'] == int(mismatch.sum())


def test_coefficient_angles_mode_and_varying_identity(small_model):
  """angles=None -> Reconstruction with each row's own angles; identity changing per frame
  (dataset-preparation callers, datasets/make_data_from_GRID.py:516-552)."""
  coeffs = synthetic.make_coeffs(6, seed=7)
  rng = np.random.Generator(np.random.PCG64(11))
  coeffs[2:, :80] = synthetic.normal_ih(rng, (4, 80)).astype(np.float32)
  coeffs[3:, 144:] = (0.1 * synthetic.normal_ih(rng, (3, 113))).astype(np.float32)
  got = render.render_sequence(coeffs, small_model, res=64, angles=None)
  want = pipeline.render_sequence(coeffs, small_model, 64, None)
  d = np.abs(got.astype(np.int16) - want.astype(np.int16))
  assert np.percentile(d, 99.5) <= 1 and (d > 1).mean() < 0.003


def test_small_model_against_reference_golden_frames(golden_small, small_model):
  g = golden_small
  got = render.render_sequence(g['coeffs'], small_model, res=64, angles=g['jitter'][:, 0, :])
  for t in range(4):
    want = g['frame%d_image' % t].reshape(64, 64, 3)
    d = np.abs(got[t].astype(np.int16) - want.astype(np.int16))
    assert np.percentile(d, 99.5) <= 1 and (d > 1).mean() < 0.003


def test_chunking_is_invisible(full_model, monkeypatch):
  coeffs = synthetic.make_coeffs(21, seed=2)
  a = np.array(render.render_sequence(coeffs, full_model, res=224))
  monkeypatch.setenv('VPB200_CHUNK_FRAMES', '5')
  b = np.array(render.render_sequence(coeffs, full_model, res=224))
  assert np.array_equal(a, b)


def test_render_face_drop_in(full_model):
  """render_face keeps the reference's signature, canvas geometry and jitter state."""
  render.reset_jitter()
  coeffs = synthetic.make_coeffs(2, seed=1)
  img = np.zeros((512, 512, 3), np.uint8)
  out = render.render_face(256, 256, 1.0, coeffs[0:1], img, [0, 0, 1.0, 0.0, 0.0], full_model)
  assert out.shape == (512, 512, 3) and out.dtype == np.uint8
  want = pipeline.render_frame(coeffs[0:1], full_model, orc.jitter_angle_sequence(1)[0, 0], 224)[0]
  face = out[256 - 112:256 + 112, 256 - 112:256 + 112, ::-1]
  d = np.abs(face.astype(np.int16) - want.astype(np.int16))
  assert np.percentile(d, 99.5) <= 1
  assert not out[:100].any()


def test_contact_sheet_matches_cpu_reference(small_model, tmp_path):
  """plot_bfm_coeff_seq (utils/bfm_visual.py:88-154): 2 x up to 30 frames through Reconstruction (coefficient
  angles), tiled 10 per row, predicted rows start at row 3; identity varies per frame in the real sequence."""
  from voicepuppet_b200 import bfm_visual
  t = 13
  real = synthetic.make_coeffs(t, seed=11)[None]
  real[0, 5:, :80] = synthetic.make_coeffs(1, seed=12)[0, :80]      # two identity runs
  pred = synthetic.make_coeffs(t, seed=13)[None, :, 80:144]
  sheet = bfm_visual.contact_sheet(small_model, [t], real, pred)
  assert sheet.shape == (9 * 224, 10 * 224, 3) and sheet.dtype == np.uint8
  spliced = np.concatenate([real[:, :, :80], pred, real[:, :, 144:]], axis=2)
  for seq, h_index in ((real, 0), (spliced, 3)):
    want = pipeline.render_sequence(np.ascontiguousarray(seq[0]), small_model, 224, None)
    for i in (0, 7, t - 1):
      r, c = i // 10 + h_index, i % 10
      tile = sheet[r * 224:(r + 1) * 224, c * 224:(c + 1) * 224]
      d = np.abs(tile[:, :, ::-1].astype(np.int16) - want[i].astype(np.int16))
      assert np.percentile(d, 99.5) <= 1 and (d > 1).mean() < 0.003
  assert not sheet[6 * 224:].any() and not sheet[2 * 224:3 * 224].any()
  bfm_visual.plot_bfm_coeff_seq(str(tmp_path), small_model, 1000, [t], real, pred)
  assert (tmp_path / 'bfmnet_1000.jpg').stat().st_size > 1000


def test_host_output_pipeline_renders_the_same_frames_as_device_outputs(full_model):
  """The host-output path (chunks drained over PCIe under the rendering of the next chunk) and the device-output
  path (two chunks in flight on two streams) cut the sequence differently; the frames must be the same."""
  import torch
  for t in (16, 75, 130):
    coeffs = synthetic.make_coeffs(t, seed=5)
    host, mask = render.render_sequence(coeffs, full_model, res=256, want_mask=True)            # host outputs: the pipeline
    dev = torch.empty((t, 256, 256, 3), dtype=torch.uint8, device='cuda:0')
    render.render_sequence(coeffs, full_model, res=256, out=dev)                                # device outputs: ChunkRunner
    torch.cuda.synchronize()
    assert np.array_equal(np.asarray(host), dev.cpu().numpy()), t
    assert mask.any()


@pytest.mark.parametrize('res,frames', [(224, 40), (256, 75), (512, 9), (1024, 5), (66, 12)])
def test_fused_kernel_is_bit_identical_to_the_separate_kernels(full_model, res, frames):
  """csrc/fused.cu (vertex stage + z-buffer scatter in one kernel, colours resolved from per-vertex colours) against
  vertex records -> scatter -> resolve: same arithmetic, so frames AND masks must be equal byte for byte, at 1-pixel
  boxes (224 / 256) as well as 17-pixel ones (1024), host-output and device-output chunking alike."""
  import torch
  from voicepuppet_b200.model import DeviceModel
  dm = DeviceModel.of(full_model)
  lib = _lib.lib()
  assert lib.vp_model_fused_available(dm.handle) == 1
  coeffs = synthetic.make_coeffs(frames, seed=11)
  dev = torch.empty((frames, res, res, 3), dtype=torch.uint8, device='cuda:0')
  _lib.check(lib.vp_set_raster_path(dm.handle, 2))
  try:
    fused, fused_mask = render.render_sequence(coeffs, full_model, res=res, want_mask=True)
    fused, fused_mask = np.asarray(fused).copy(), np.asarray(fused_mask).copy()
    render.render_sequence(coeffs, full_model, res=res, out=dev)
    torch.cuda.synchronize()
  finally:
    _lib.check(lib.vp_set_raster_path(dm.handle, 0))
  sep, sep_mask = render.render_sequence(coeffs, full_model, res=res, want_mask=True)
  sep, sep_mask = np.asarray(sep).copy(), np.asarray(sep_mask).copy()
  assert fused_mask.any() and np.array_equal(fused_mask, sep_mask)
  assert np.array_equal(fused, sep)
  assert np.array_equal(dev.cpu().numpy(), sep)


def test_fused_kernel_small_model_and_coefficient_angles(small_model):
  """Same on the 420-vertex model (tiles with few triangles, partial warps) and with Reconstruction's single
  rotation from the coefficient rows (angles=None)."""
  from voicepuppet_b200.model import DeviceModel
  dm = DeviceModel.of(small_model)
  lib = _lib.lib()
  assert lib.vp_model_fused_available(dm.handle) == 1
  coeffs = synthetic.make_coeffs(7, seed=3)
  for angles in ('jitter', None):
    a = np.asarray(render.render_sequence(coeffs, small_model, res=96, angles=angles)).copy()
    _lib.check(lib.vp_set_raster_path(dm.handle, 2))
    try:
      b = np.asarray(render.render_sequence(coeffs, small_model, res=96, angles=angles)).copy()
    finally:
      _lib.check(lib.vp_set_raster_path(dm.handle, 0))
    assert a.any() and np.array_equal(a, b)
