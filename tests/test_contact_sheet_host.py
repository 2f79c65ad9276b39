"""Host logic of the contact-sheet drop-in (voicepuppet_b200/bfm_visual.py) against the LIVE reference function
utils/bfm_visual.py:88-154 (where /root/reference exists): same splicing of identity / texture / expression
columns, same tiling, same channel order.  The renderer is replaced by the CPU oracle so that only the host logic
is compared -- exactly; the GPU renderer itself is covered by tests/test_gpu_sequence.py."""
import os
import sys

import numpy as np
import pytest

from oracle import build_ref, pipeline
from voicepuppet_b200 import bfm_visual
from voicepuppet_b200 import synthetic

REFERENCE_UTILS = '/root/reference/utils'


def oracle_render(coeffs, facemodel, res=224, angles=None):
  return pipeline.render_sequence(coeffs, facemodel, res, angles)


@pytest.fixture(scope='module')
def live_plot():
  if not os.path.exists(os.path.join(REFERENCE_UTILS, 'bfm_visual.py')):
    pytest.skip('/root/reference not present on this host')
  pytest.importorskip('cv2')
  if not build_ref.build() or not os.path.exists(build_ref.cython_module_path()):
    pytest.skip('the reference Cython module could not be built')
  sys.path.insert(0, os.path.dirname(build_ref.cython_module_path()))
  sys.path.insert(0, REFERENCE_UTILS)
  cwd = os.getcwd()
  try:
    import bfm_visual as ref_visual          # utils/bfm_visual.py, unmodified
  except Exception as e:                      # optional third-party imports of the reference module
    pytest.skip('reference bfm_visual not importable here: %s' % e)
  finally:
    os.chdir(cwd)
  return ref_visual


@pytest.mark.parametrize('with_identity', [False, True])
def test_contact_sheet_equals_live_reference(live_plot, small_model, tmp_path, monkeypatch, with_identity):
  t = 12
  real = synthetic.make_coeffs(t, seed=31)[None]
  real[0, 4:, :80] = synthetic.make_coeffs(1, seed=32)[0, :80]
  pred = synthetic.make_coeffs(t, seed=33)[None, :, 80:144]
  idc = synthetic.make_coeffs(1, seed=34)[None, :, :80] if with_identity else None
  texc = synthetic.make_coeffs(1, seed=35)[None, :, 144:224] if with_identity else None
  captured = {}
  monkeypatch.setattr(live_plot.cv2, 'imwrite', lambda path, img: captured.update(path=path, img=img.copy()) or True)
  live_plot.plot_bfm_coeff_seq(str(tmp_path), small_model, 7, [t], real, pred, idc, texc)
  assert captured['path'] == '{}/bfmnet_{}.jpg'.format(str(tmp_path), 7)
  ours = bfm_visual.contact_sheet(small_model, [t], real, pred, idc, texc, render_fn=oracle_render)
  assert ours.shape == captured['img'].shape and np.array_equal(ours, captured['img'])


def test_sequences_longer_than_30_are_trimmed(small_model):
  calls = []

  def fake(coeffs, facemodel, res=224, angles=None):
    calls.append(coeffs.shape[0])
    return np.zeros((coeffs.shape[0], res, res, 3), np.uint8)

  real = np.zeros((1, 41, 257), np.float32)
  sheet = bfm_visual.contact_sheet(small_model, [41], real, real[:, :, 80:144], render_fn=fake)
  assert calls == [30, 30] and sheet.shape == (9 * 224, 10 * 224, 3)
