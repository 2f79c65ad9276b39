"""Pins oracle/composite.py (restatement of cv2.resize's 8-bit bilinear path + the paste / float conversion of
infer_bfmvid.py:111-121,234) against golden outputs made with cv2 in this container and, where cv2 is
importable, against cv2 itself on random sizes."""
import os

import numpy as np
import pytest

from oracle import composite as oc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'composite.npz')


@pytest.fixture(scope='module')
def golden():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def test_composite_matches_reference_golden(golden):
  for i in range(int(golden['n_cases'])):
    a = golden['c%d_args' % i]
    canvas, face3d = oc.composite(golden['c%d_raster' % i], int(a[0]), int(a[1]), float(a[2]), a[3:], (512, 512))
    assert np.array_equal(canvas, golden['c%d_canvas' % i]), i
    assert face3d.dtype == np.float32
    assert np.array_equal(face3d, canvas[:, :, ::-1].astype(np.float32) / 255.0)
    assert float(face3d.astype(np.float64).sum()) == float(golden['c%d_face3d_sum' % i])


@pytest.mark.parametrize('seed', range(4))
def test_resize_matches_cv2(seed):
  cv2 = pytest.importorskip('cv2')
  rng = np.random.Generator(np.random.PCG64(seed))
  for _ in range(6):
    sh, sw = int(rng.integers(16, 300)), int(rng.integers(16, 300))
    dh, dw = int(rng.integers(8, 500)), int(rng.integers(8, 500))
    src = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
    assert np.array_equal(oc.resize_linear_u8(src, dw, dh), cv2.resize(src, (dw, dh))), (sh, sw, dh, dw)
  src = rng.integers(0, 256, (224, 224, 3)).astype(np.uint8)
  for s in (112, 224, 223, 225, 448, 100):
    assert np.array_equal(oc.resize_linear_u8(src, s, s), cv2.resize(src, (s, s))), s


def test_paste_outside_the_canvas_raises_like_numpy():
  raster = np.zeros((224, 224, 3), np.uint8)
  with pytest.raises(ValueError):
    oc.composite(raster, 500, 256, 1.0, np.array([512, 512, 1.0, 0, 0.0]), (512, 512))


def test_c_abi_placement_matches_the_reference_lines():
  """vp_composite_placement (host only, no GPU): infer_bfmvid.py:80-82,112-121 -- Python's int() truncation
  and round-half-even -- against the oracle on random parameters, including exact .5 sizes."""
  from voicepuppet_b200 import render
  rng = np.random.Generator(np.random.PCG64(9))
  cases = [(256, 250, 1.05, [512, 512, 0.97, 12.3, -20.7]), (100, 100, 224 / 150.5, [0, 0, 1.0, 0.0, 0.0]),
           (100, 100, 224 / 151.5, [0, 0, 1.0, -0.49, 0.51])]
  for _ in range(200):
    cases.append((int(rng.integers(0, 600)), int(rng.integers(0, 600)), float(rng.uniform(0.4, 2.5)),
                  [512, 512, float(rng.uniform(0.7, 1.4)), float(rng.uniform(-60, 60)), float(rng.uniform(-60, 60))]))
  for cx, cy, ratio, tp in cases:
    for res in (224, 256):
      assert render.composite_placement(cx, cy, ratio, tp, res) == oc.placement(cx, cy, ratio, np.array(tp), res), (cx, cy, ratio, tp, res)


def test_c_abi_axis_tables_match_the_cv2_pinned_oracle():
  """The C++ builder of the resize coefficient tables (host code of csrc/composite.cu) over many size pairs."""
  from voicepuppet_b200 import _lib
  lib = _lib.lib()
  rng = np.random.Generator(np.random.PCG64(17))
  pairs = [(224, s) for s in range(60, 460, 7)] + [(int(a), int(b)) for a, b in rng.integers(8, 700, (120, 2))]
  for ssize, dsize in pairs:
    for is_y in (0, 1):
      out = np.zeros((dsize, 4), np.int32)
      _lib.check(lib.vp_composite_axis_table(ssize, dsize, is_y, _lib.ptr(out)))
      s0, s1, c0, c1 = oc.axis_table(ssize, dsize, bool(is_y))
      assert np.array_equal(out, np.stack([s0, s1, c0, c1], axis=1)), (ssize, dsize, is_y)
