"""GPU parity for the on-device post-raster composite: bit-exact against golden outputs made
with the reference's own cv2 / numpy lines (tests/golden/make_golden_composite.py) and against the oracle.
Reference: voicepuppet/pixrefer/infer_bfmvid.py:79-82, 111-121, 234-236."""
import os

import numpy as np
import pytest

from oracle import composite as oc, pipeline, reconstruct_oracle as orc
from voicepuppet_b200 import _lib, render, synthetic

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'composite.npz')


@pytest.fixture(scope='module')
def golden():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def test_composite_matches_reference_golden(golden):
  import torch
  dev = torch.device('cuda', 0)
  for i in range(int(golden['n_cases'])):
    a = golden['c%d_args' % i]
    raster = torch.from_numpy(golden['c%d_raster' % i][None]).to(dev)
    inputs = torch.full((1, 512, 512, 6), -7.0, dtype=torch.float32, device=dev)
    canvas, _ = render.composite_device(raster, int(a[0]), int(a[1]), float(a[2]), a[3:], (512, 512), inputs, 3)
    want = golden['c%d_canvas' % i]
    assert np.array_equal(canvas[0].cpu().numpy(), want), i
    got_in = inputs[0].cpu().numpy()
    assert np.all(got_in[..., 0:3] == -7.0)                                   # other channels untouched
    assert np.array_equal(got_in[..., 3:6], want[:, :, ::-1].astype(np.float32) / 255.0)


@pytest.mark.parametrize('seed', range(3))
def test_random_batches_match_oracle(seed):
  import torch
  rng = np.random.Generator(np.random.PCG64(700 + seed))
  dev = torch.device('cuda', 0)
  res = [224, 256, 96][seed]
  t = 5
  rasters = rng.integers(0, 256, (t, res, res, 3)).astype(np.uint8)
  ratio = [1.13, 0.77, 2.0][seed]
  tp = np.array([512, 512, [0.93, 1.0, 1.0][seed], 17.2, -9.6])
  hw = (600, 640)
  canvas, inputs = render.composite_device(torch.from_numpy(rasters).to(dev), 300, 290, ratio, tp, hw,
                                           torch.zeros((t, hw[0], hw[1], 3), dtype=torch.float32, device=dev), 0)
  for k in range(t):
    wc, wf = oc.composite(rasters[k], 300, 290, ratio, tp, hw)
    assert np.array_equal(canvas[k].cpu().numpy(), wc)
    assert np.array_equal(inputs[k].cpu().numpy(), wf)


def test_face_outside_the_canvas_is_an_error():
  import torch
  raster = torch.zeros((1, 224, 224, 3), dtype=torch.uint8, device='cuda:0')
  with pytest.raises(_lib.VpError):
    render.composite_device(raster, 500, 256, 1.0, [512, 512, 1.0, 0.0, 0.0], (512, 512))


def test_render_face_sequence_feeds_the_network_input(full_model):
  """Coefficients -> PixReferNet input channels 3:6 without leaving the GPU, against the CPU reference path
  (oracle pipeline + oracle composite); the rasterizer's own tolerance applies (edge pixels, see
  test_gpu_sequence), everything after it is exact."""
  import torch
  t = 6
  coeffs = synthetic.make_coeffs(t, seed=1)
  tp = np.array([512, 512, 0.96, 10.0, -14.0])
  inputs = torch.zeros((t, 512, 512, 6), dtype=torch.float32, device='cuda:0')
  canvas, _ = render.render_face_sequence(256, 250, 1.1, coeffs, (512, 512, 3), tp, full_model, inputs)
  jit = orc.jitter_angle_sequence(t)
  for k in (0, t - 1):
    raster = pipeline.render_frame(coeffs[k:k + 1], full_model, jit[k, 0], 224)[0]
    wc, wf = oc.composite(raster, 256, 250, 1.1, tp, (512, 512))
    d = np.abs(canvas[k].cpu().numpy().astype(np.int16) - wc.astype(np.int16))
    assert np.percentile(d, 99.9) <= 1 and (d > 1).mean() < 1e-3
    got = inputs[k, ..., 3:6].cpu().numpy()
    assert np.array_equal(got, canvas[k].cpu().numpy()[:, :, ::-1].astype(np.float32) / 255.0)
    assert not inputs[k, ..., 0:3].any()
