"""TEST INFRASTRUCTURE -- Python doorways onto the two CPU rasterizers.

  * ``Oracle``     : oracle/mesh_core_oracle.c (our order-independent restatement)
  * ``Reference``  : oracle/_ref/libmesh_core_ref.so, the reference's own
                     utils/cython/mesh_core.cpp compiled unmodified (oracle/build_ref.py)

Both take the reference's flat buffers (utils/cython/mesh_core_cython.pyx:49-78) and
mutate them in place.
"""
import ctypes
import os

import numpy as np

from . import build_oracle, build_ref

_u8p = ctypes.POINTER(ctypes.c_ubyte)
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def _p(a, ty):
  return a.ctypes.data_as(ty)


def _chk(a, dtype):
  assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous, (a.dtype, dtype)
  return a


class Oracle(object):
  """Order-independent C restatement."""
  _lib = None

  @classmethod
  def lib(cls):
    if cls._lib is None:
      lib = ctypes.CDLL(build_oracle.build())
      lib.vpo_render_colors.argtypes = [_u8p, _u8p, _f32p, _i32p, _f32p, _f32p, _i32p] + [ctypes.c_int] * 5
      lib.vpo_render_colors.restype = ctypes.c_int
      lib.vpo_rasterize_triangles.argtypes = [_f32p, _i32p, _f32p, _i32p, _f32p] + [ctypes.c_int] * 5
      lib.vpo_rasterize_triangles.restype = ctypes.c_int
      lib.vpo_render_colors_near_ties.argtypes = [_f32p, _i32p] + [ctypes.c_int] * 4 + [_u8p]
      lib.vpo_render_colors_near_ties.restype = ctypes.c_int
      lib.vpo_render_texture.argtypes = [_f32p, _f32p, _i32p, _f32p, _f32p, _i32p, _f32p] + [ctypes.c_int] * 11
      lib.vpo_render_texture.restype = ctypes.c_int
      lib.vpo_get_normal.argtypes = [_f32p, _f32p, _i32p, ctypes.c_int, ctypes.c_int]
      lib.vpo_get_normal.restype = ctypes.c_int
      cls._lib = lib
    return cls._lib

  @classmethod
  def render_colors(cls, image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c,
                    triangle_out=None, reverse=False):
    tri_ptr = _p(_chk(triangle_out, np.int32), _i32p) if triangle_out is not None else None
    rc = cls.lib().vpo_render_colors(
        _p(_chk(image, np.uint8), _u8p), _p(_chk(face_mask, np.uint8), _u8p),
        _p(_chk(vertices, np.float32), _f32p), _p(_chk(triangles, np.int32), _i32p),
        _p(_chk(colors, np.float32), _f32p), _p(_chk(depth_buffer, np.float32), _f32p),
        tri_ptr, ntri, h, w, c, int(reverse))
    assert rc == 0

  @classmethod
  def rasterize_triangles(cls, vertices, triangles, depth_buffer, triangle_buffer, barycentric_weight,
                          nver, ntri, h, w, reverse=False):
    rc = cls.lib().vpo_rasterize_triangles(
        _p(_chk(vertices, np.float32), _f32p), _p(_chk(triangles, np.int32), _i32p),
        _p(_chk(depth_buffer, np.float32), _f32p), _p(_chk(triangle_buffer, np.int32), _i32p),
        _p(_chk(barycentric_weight, np.float32), _f32p), nver, ntri, h, w, int(reverse))
    assert rc == 0

  @classmethod
  def render_texture(cls, image, vertices, triangles, texture, tex_coords, tex_triangles, depth_buffer,
                     nver, tex_nver, ntri, h, w, c, tex_h, tex_w, tex_c, mapping_type, reverse=False):
    rc = cls.lib().vpo_render_texture(
        _p(_chk(image, np.float32), _f32p), _p(_chk(vertices, np.float32), _f32p),
        _p(_chk(triangles, np.int32), _i32p), _p(_chk(texture, np.float32), _f32p),
        _p(_chk(tex_coords, np.float32), _f32p), _p(_chk(tex_triangles, np.int32), _i32p),
        _p(_chk(depth_buffer, np.float32), _f32p),
        nver, tex_nver, ntri, h, w, c, tex_h, tex_w, tex_c, mapping_type, int(reverse))
    assert rc == 0

  @classmethod
  def get_normal(cls, normal, tri_normal, triangles, ntri):
    rc = cls.lib().vpo_get_normal(_p(_chk(normal, np.float32), _f32p), _p(_chk(tri_normal, np.float32), _f32p),
                                  _p(_chk(triangles, np.int32), _i32p), normal.size // 3, ntri)
    assert rc == 0

  @classmethod
  def near_ties(cls, vertices, triangles, ntri, h, w, ulps=1):
    out = np.zeros(h * w, dtype=np.uint8)
    n = cls.lib().vpo_render_colors_near_ties(
        _p(_chk(vertices, np.float32), _f32p), _p(_chk(triangles, np.int32), _i32p), ntri, h, w, ulps,
        _p(out, _u8p))
    assert n >= 0
    return out


class Reference(object):
  """The reference's own C++ (sequential loops), via oracle/ref_shim.cpp."""
  _lib = None

  @classmethod
  def available(cls):
    return build_ref.build() and os.path.exists(build_ref.lib_path())

  @classmethod
  def lib(cls):
    if cls._lib is None:
      if not cls.available():
        raise RuntimeError('oracle/_ref/libmesh_core_ref.so is missing (run oracle/build_ref.py '
                           'where /root/reference exists)')
      lib = ctypes.CDLL(build_ref.lib_path())
      lib.ref_render_colors_core.argtypes = [_u8p, _u8p, _f32p, _i32p, _f32p, _f32p] + [ctypes.c_int] * 4
      lib.ref_render_colors_core.restype = None
      lib.ref_rasterize_triangles_core.argtypes = [_f32p, _i32p, _f32p, _i32p, _f32p] + [ctypes.c_int] * 4
      lib.ref_rasterize_triangles_core.restype = None
      lib.ref_render_texture_core.argtypes = [_f32p, _f32p, _i32p, _f32p, _f32p, _i32p, _f32p] + [ctypes.c_int] * 10
      lib.ref_render_texture_core.restype = None
      lib.ref_get_normal_core.argtypes = [_f32p, _f32p, _i32p, ctypes.c_int]
      lib.ref_get_normal_core.restype = None
      cls._lib = lib
    return cls._lib

  @classmethod
  def render_colors(cls, image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c):
    cls.lib().ref_render_colors_core(
        _p(_chk(image, np.uint8), _u8p), _p(_chk(face_mask, np.uint8), _u8p),
        _p(_chk(vertices, np.float32), _f32p), _p(_chk(triangles, np.int32), _i32p),
        _p(_chk(colors, np.float32), _f32p), _p(_chk(depth_buffer, np.float32), _f32p), ntri, h, w, c)

  @classmethod
  def rasterize_triangles(cls, vertices, triangles, depth_buffer, triangle_buffer, barycentric_weight,
                          nver, ntri, h, w):
    cls.lib().ref_rasterize_triangles_core(
        _p(_chk(vertices, np.float32), _f32p), _p(_chk(triangles, np.int32), _i32p),
        _p(_chk(depth_buffer, np.float32), _f32p), _p(_chk(triangle_buffer, np.int32), _i32p),
        _p(_chk(barycentric_weight, np.float32), _f32p), nver, ntri, h, w)

  @classmethod
  def render_texture(cls, image, vertices, triangles, texture, tex_coords, tex_triangles, depth_buffer,
                     nver, tex_nver, ntri, h, w, c, tex_h, tex_w, tex_c, mapping_type):
    cls.lib().ref_render_texture_core(
        _p(_chk(image, np.float32), _f32p), _p(_chk(vertices, np.float32), _f32p),
        _p(_chk(triangles, np.int32), _i32p), _p(_chk(texture, np.float32), _f32p),
        _p(_chk(tex_coords, np.float32), _f32p), _p(_chk(tex_triangles, np.int32), _i32p),
        _p(_chk(depth_buffer, np.float32), _f32p),
        nver, tex_nver, ntri, h, w, c, tex_h, tex_w, tex_c, mapping_type)

  @classmethod
  def get_normal(cls, normal, tri_normal, triangles, ntri):
    cls.lib().ref_get_normal_core(_p(_chk(normal, np.float32), _f32p), _p(_chk(tri_normal, np.float32), _f32p),
                                  _p(_chk(triangles, np.int32), _i32p), ntri)


def fresh_color_buffers(h, w, c=3):
  """The buffers infer_bfmvid.render_face allocates (voicepuppet/pixrefer/infer_bfmvid.py:100-106)."""
  image = np.zeros(h * w * c, dtype=np.uint8)
  mask = np.zeros(h * w, dtype=np.uint8)
  depth = (np.zeros(h * w) - 99999.0).astype(np.float32)
  return image, mask, depth
