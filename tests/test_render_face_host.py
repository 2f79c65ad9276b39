"""Host logic of the three render_face flavours (voicepuppet/pixrefer/infer_bfmvid.py:79-122,
voicepuppet/pixflow/infer_bfm_pixflow.py:72-115, datasets/make_data_from_GRID.py:516-552): which angles reach the
renderer and how the jitter state advances.  The GPU calls are replaced by recorders (no device needed)."""
import sys
import types

import numpy as np
import pytest

from oracle import reconstruct_oracle as orc
from voicepuppet_b200 import render


@pytest.fixture()
def recorded(monkeypatch):
  calls = []

  class FakeTensor(object):
    def __init__(self, shape):
      self.shape = shape

    def __getitem__(self, i):
      return self

    def cpu(self):
      return self

    def numpy(self):
      return np.zeros((512, 512, 3), np.uint8)

  fake_torch = types.SimpleNamespace(
      device=lambda *a: ('cuda', 0), uint8='u8',
      empty=lambda shape, dtype=None, device=None: FakeTensor(shape),
      cuda=types.SimpleNamespace(device=lambda d: __import__('contextlib').nullcontext()))
  monkeypatch.setitem(sys.modules, 'torch', fake_torch)
  monkeypatch.setattr(render, 'render_sequence',
                      lambda coeffs, model, res, angles, device, out: calls.append(None if angles is None else np.array(angles, copy=True)))
  monkeypatch.setattr(render, 'composite_device', lambda frames, cx, cy, ratio, tp, hw: (FakeTensor((1,) + tuple(hw) + (3,)), None))
  render.reset_jitter()
  return calls


def test_flavours_pass_the_reference_angles(recorded):
  img = np.zeros((512, 512, 3), np.uint8)
  c = np.zeros((1, 257), np.float32)
  for _ in range(3):
    out = render.render_face(256, 256, 1.0, c, img, [0, 0, 1.0, 0, 0], None)
    assert out.shape == img.shape and out.dtype == img.dtype
  want = orc.jitter_angle_sequence(3)
  for k in range(3):
    assert np.array_equal(recorded[k].reshape(1, 3), want[k])               # infer_bfmvid.py:85-89
  render.render_face_pixflow(256, 256, 1.0, c, img, [0, 0, 1.0, 0, 0], None)
  assert np.array_equal(recorded[3], np.zeros((1, 3), np.float32))          # jitter commented out in the pixflow copy
  render.render_face_dataset(256, 256, 1.0, c, img, [0, 0, 1.0, 0, 0], None)
  assert recorded[4] is None                                                # Reconstruction with the row's own angles
  render.render_face(256, 256, 1.0, c, img, [0, 0, 1.0, 0, 0], None)
  assert np.array_equal(recorded[5].reshape(1, 3), orc.jitter_angle_sequence(4)[3])   # the other flavours left the state alone
