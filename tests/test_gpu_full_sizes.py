"""BASELINE.json's full-size configurations, checked through size-independent properties (the CPU oracle
takes ~70 ms per frame, so whole-sequence comparisons are made on samples):

  config 2  1500 frames at 512x512 on one GPU          (1-minute clip)
  config 4  4096 frames at 1024x1024, tcgen05 basis    (stress; 12.9 GB of frames, kept on the device)
  config 1  75 frames at 256x256                        covered frame by frame in test_gpu_sequence.py
  config 3  12000 frames over 2/4/8 GPUs                tests/test_gpu_sharded.py (2 GPUs) + bench.py --gpus N

Properties:
  * periodicity: the jitter sequence (infer_bfmvid.py:85-89) has period 28, so with expression coefficients
    tiled with period 28 frame t and frame t + 28k are bit-identical, whatever chunk / basis group / stream
    they were rendered in;
  * chunk independence: a frame rendered inside the long sequence equals the same frame rendered alone;
  * oracle samples: a few frames against the CPU reference path at the full resolution;
  * a checksum of per-frame checksums equal between two renderings with different chunking."""
import numpy as np
import pytest

from oracle import pipeline, reconstruct_oracle as orc
from voicepuppet_b200 import render, synthetic

pytestmark = pytest.mark.gpu
PERIOD = 28


def periodic_coeffs(t, seed):
  base = synthetic.make_coeffs(PERIOD, seed=seed)
  reps = -(-t // PERIOD)
  return np.ascontiguousarray(np.tile(base, (reps, 1))[:t])


def frame_checksums(frames):
  """uint64 per frame, on the device: sum of bytes weighted by a position-dependent odd multiplier."""
  import torch
  t = frames.shape[0]
  flat = frames.reshape(t, -1).to(torch.int64)
  w = (torch.arange(flat.shape[1], device=frames.device, dtype=torch.int64) * 2654435761 + 12345) % 1000003
  return (flat * w).sum(dim=1)


def test_jitter_sequence_has_period_28():
  a = render.jitter_angle_sequence(3 * PERIOD + 5)
  assert np.array_equal(a[:PERIOD + 5], a[PERIOD:2 * PERIOD + 5]) and np.array_equal(a[:PERIOD], a[2 * PERIOD:3 * PERIOD])


def test_one_minute_clip_1500_frames_at_512(full_model, monkeypatch):
  import torch
  t, res = 1500, 512
  coeffs = periodic_coeffs(t, seed=21)
  out = torch.empty((t, res, res, 3), dtype=torch.uint8, device='cuda:0')
  render.render_sequence(coeffs, full_model, res=res, angles='jitter', out=out)
  torch.cuda.synchronize()
  sums = frame_checksums(out).cpu().numpy()
  # periodicity across chunks, basis groups and streams
  assert np.array_equal(sums[:t - PERIOD], sums[PERIOD:])
  assert len(set(sums[:PERIOD].tolist())) == PERIOD            # and the frames of one period do differ
  assert torch.equal(out[3], out[3 + 28 * 40]) and torch.equal(out[27], out[27 + 28 * 52])
  # chunk independence: other chunking, same bytes (checksum of checksums)
  monkeypatch.setenv('VPB200_CHUNK_FRAMES', '37')
  out2 = torch.empty_like(out)
  render.render_sequence(coeffs, full_model, res=res, angles='jitter', out=out2)
  torch.cuda.synchronize()
  monkeypatch.delenv('VPB200_CHUNK_FRAMES')
  sums2 = frame_checksums(out2).cpu().numpy()
  assert int(sums.sum()) == int(sums2.sum()) and np.array_equal(sums, sums2)
  # oracle samples at full resolution (first period; the rest follows by periodicity)
  jit = orc.jitter_angle_sequence(PERIOD)[:, 0, :]
  for k in (0, 13, 27):
    want = pipeline.render_frame(coeffs[k:k + 1], full_model, jit[k], res)[0]
    d = np.abs(out[k].cpu().numpy().astype(np.int16) - want.astype(np.int16)).max(axis=2)
    assert np.percentile(d, 99.9) <= 1 and (d > 1).sum() <= 40, (k, int((d > 1).sum()))
  # a frame rendered alone equals the frame inside the sequence (same basis kernel: a single frame would
  # otherwise take the FP32 SIMT flavour, whose last-bit differences can move an edge pixel)
  from voicepuppet_b200 import _lib
  from voicepuppet_b200.model import DeviceModel
  dm = DeviceModel.of(full_model)
  try:
    _lib.check(_lib.lib().vp_set_basis_mode(dm.handle, 2))
    alone = np.asarray(render.render_sequence(coeffs[700:701], full_model, res=res,
                                              angles=render.jitter_angle_sequence(701)[700:701]))
  finally:
    _lib.check(_lib.lib().vp_set_basis_mode(dm.handle, 0))
  assert np.array_equal(alone[0], out[700].cpu().numpy())


def test_stress_4096_frames_at_1024_tensor_core_basis(full_model):
  import torch
  from voicepuppet_b200 import _lib
  from voicepuppet_b200.model import DeviceModel
  t, res = 4096, 1024
  free, _ = torch.cuda.mem_get_info(0)
  if free < 40 * (1 << 30):
    pytest.skip('needs ~14 GB for the frames plus workspaces')
  dm = DeviceModel.of(full_model)
  coeffs = periodic_coeffs(t, seed=22)
  out = torch.empty((t, res, res, 3), dtype=torch.uint8, device='cuda:0')
  try:
    _lib.check(_lib.lib().vp_set_basis_mode(dm.handle, 2))       # tcgen05 3xTF32 forced on (it is the default here anyway)
    render.render_sequence(coeffs, full_model, res=res, angles='jitter', out=out)
    torch.cuda.synchronize()
  finally:
    _lib.check(_lib.lib().vp_set_basis_mode(dm.handle, 0))
  sums = torch.stack([frame_checksums(out[a:a + 256]) for a in range(0, t, 256)]).reshape(-1).cpu().numpy()
  assert np.array_equal(sums[:t - PERIOD], sums[PERIOD:])
  assert len(set(sums[:PERIOD].tolist())) == PERIOD
  jit = orc.jitter_angle_sequence(PERIOD)[:, 0, :]
  k = 5
  want = pipeline.render_frame(coeffs[k:k + 1], full_model, jit[k], res)[0]
  got = out[k + 28 * 100].cpu().numpy()
  d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max(axis=2)
  assert np.percentile(d, 99.9) <= 1 and (d > 1).sum() <= 120, int((d > 1).sum())
  assert (got.max(axis=2) > 0).mean() > 0.4                      # ~47 % coverage at every size
  del out
  torch.cuda.empty_cache()


_WALK_SCRIPT = r'''
import hashlib, sys
import numpy as np
from voicepuppet_b200 import render, synthetic
full = synthetic.cached_model()
coeffs = synthetic.make_coeffs(70, seed=31)
frames = np.asarray(render.render_sequence(coeffs, full, res=int(sys.argv[1]), angles='jitter'))
print('SHA', hashlib.sha1(frames.tobytes()).hexdigest(), int(frames.any()))
'''


@pytest.mark.parametrize('res', [768, 1024])
def test_group_walk_equals_lane_local_walk(res):
  """From 768x768 the scatter kernel walks the boxes of four lanes together (raster_walk.cuh); the walk is chosen once per
  process, so two child processes render the same 70 frames, one with the group walk (default) and one with every lane
  walking its own box (VPB200_WALK_GROUP=0): the frames must be the same bytes.  (Each walk is also compared with the
  CPU reference: test_stress_4096_frames_at_1024_tensor_core_basis samples frames of the default, test_gpu_sequence of the lane-local one.)"""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  got = {}
  for group in ('4', '0'):
    env = dict(os.environ, VPB200_WALK_GROUP=group, PYTHONPATH=root + os.pathsep + os.environ.get('PYTHONPATH', ''))
    out = subprocess.run([sys.executable, '-c', _WALK_SCRIPT, str(res)], env=env, cwd=root, capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('SHA')][-1].split()
    assert line[2] == '1'
    got[group] = line[1]
  assert got['4'] == got['0']
