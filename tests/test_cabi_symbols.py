"""The C-ABI shared library loads without a GPU and exports every function include/vpb200.h
declares; the ctypes table in voicepuppet_b200/_lib.py covers exactly that set."""
import ctypes
import os
import re

import pytest

from voicepuppet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
  text = open(os.path.join(ROOT, 'include', 'vpb200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(vp_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_path():
  names = declared_functions()
  for must in ('vp_render_colors_core', 'vp_rasterize_triangles_core', 'vp_model_create', 'vp_set_identity',
               'vp_reconstruct', 'vp_render_sequence', 'vp_render_sequence_dev', 'vp_last_error'):
    assert must in names


def test_library_exports_every_declared_symbol():
  lib = ctypes.CDLL(_lib.LIB_PATH)
  missing = [n for n in declared_functions() if not hasattr(lib, n)]
  assert not missing, missing


def test_ctypes_table_matches_header():
  assert sorted(_lib.SIGNATURES) == declared_functions()


def test_no_cpu_fallback_without_a_device():
  lib = _lib.lib()
  if lib.vp_device_count() > 0:
    pytest.skip('a CUDA device is present')
  import numpy as np
  from voicepuppet_b200 import mesh_core_cython
  image = np.zeros(8 * 8 * 3, np.uint8)
  mask = np.zeros(64, np.uint8)
  depth = np.full(64, -99999.0, np.float32)
  verts = np.array([0, 0, 1, 5, 0, 1, 0, 5, 1], np.float32)
  with pytest.raises(_lib.VpError):
    mesh_core_cython.render_colors_core(image, mask, verts, np.array([0, 1, 2], np.int32), verts.copy(), depth, 1, 8, 8, 3)
  assert not image.any()


def test_cython_style_argument_checks():
  import numpy as np
  from voicepuppet_b200 import mesh_core_cython as mc
  f32 = np.zeros(9, np.float32)
  u8 = np.zeros(64 * 3, np.uint8)
  with pytest.raises(TypeError):
    mc.render_colors_core(None, u8, f32, np.zeros(3, np.int32), f32, np.zeros(64, np.float32), 1, 8, 8, 3)
  with pytest.raises(ValueError):   # wrong dtype
    mc.render_colors_core(u8, u8, f32.astype(np.float64), np.zeros(3, np.int32), f32, np.zeros(64, np.float32), 1, 8, 8, 3)
  with pytest.raises(ValueError):   # wrong ndim
    mc.render_colors_core(u8, u8, f32.reshape(3, 3), np.zeros(3, np.int32), f32, np.zeros(64, np.float32), 1, 8, 8, 3)
  with pytest.raises(ValueError):   # not contiguous
    mc.render_colors_core(u8, u8, np.zeros(18, np.float32)[::2], np.zeros(3, np.int32), f32, np.zeros(64, np.float32), 1, 8, 8, 3)
