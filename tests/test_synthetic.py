"""Host logic: the synthetic BFM-shaped model honours the reference's model contract
(utils/bfm_load_data.py:9-21)."""
import numpy as np

from voicepuppet_b200 import synthetic


def test_full_model_contract(full_model):
  m = full_model
  n, f = synthetic.N_VERTICES, synthetic.N_TRIANGLES
  assert m.meanshape.shape == (1, 3 * n) and m.meantex.shape == (1, 3 * n)
  assert m.idBase.shape == (3 * n, 80) and m.exBase.shape == (3 * n, 64) and m.texBase.shape == (3 * n, 80)
  assert m.tri.shape == (f, 3) and m.point_buf.shape == (n, 8) and m.keypoints.shape == (68,)
  assert m.exBase.dtype == np.float64 and m.idBase.dtype == np.float32
  tri = m.tri.astype(np.int64)
  assert tri.min() == 1 and tri.max() == n                     # 1-based
  pb = m.point_buf.astype(np.int64)
  assert pb.min() >= 1 and pb.max() == f + 1                   # pad value F+1
  # every (vertex, face) incidence appears exactly once, in triangle order
  for v in (0, 1234, n - 1):
    faces = [x for x in pb[v] if x <= f]
    assert faces == sorted(faces)
    assert all(v + 1 in tri[x - 1] for x in faces)
    assert len(faces) == int((tri == v + 1).any(axis=1).sum())


def test_small_model_reproducible(small_model):
  again = synthetic.make_model(small_model.meanshape.shape[1] // 3, 48)
  assert again.checksum() == small_model.checksum()


def test_coeff_layout():
  c = synthetic.make_coeffs(20)
  assert c.shape == (20, 257) and c.dtype == np.float32
  # only the 64 expression columns vary per frame (infer_bfmvid.py:223-224)
  assert np.all(c[:, :80] == c[0, :80]) and np.all(c[:, 144:] == c[0, 144:])
  assert not np.all(c[:, 80:144] == c[0, 80:144])
