"""Drop-in for the reference's native rasterizer module ``utils/cython/mesh_core_cython.pyx``.

    from voicepuppet_b200 import mesh_core_cython
    mesh_core_cython.render_colors_core(image, face_mask, vertices, triangles, colors, depth_buffer,
                                        ntri, h, w, c)

Same contract as the Cython wrappers (mesh_core_cython.pyx:49-78): the caller allocates and
pre-initialises every buffer, the call mutates them in place and returns None; ``None`` arguments
raise TypeError, wrong dtype / ndim / non-contiguous arrays raise ValueError like Cython's typed
buffer check.  The work is done by libvpb200.so's z-buffer kernels (order-independent 64-bit
(depth, triangle) atomicMax + resolve), bit-identical to the reference's sequential C++ loop
(utils/cython/mesh_core.cpp:108-231).
"""
import numpy as np

from . import _lib

__all__ = ['render_colors_core', 'rasterize_triangles_core', 'render_texture_core', 'get_normal_core']


def _buffer(name, a, dtype, ndim):
  if a is None:
    raise TypeError("Argument '%s' must not be None" % name)
  if not isinstance(a, np.ndarray):
    raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)" % (name, type(a).__name__))
  if a.dtype != dtype:
    raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % (np.dtype(dtype).name, a.dtype.name))
  if a.ndim != ndim:
    raise ValueError('Buffer has wrong number of dimensions (expected %d, got %d)' % (ndim, a.ndim))
  if not a.flags.c_contiguous:
    raise ValueError('ndarray is not C-contiguous')
  return a


def _need(name, a, count):
  if a.size < count:
    raise ValueError("buffer '%s' holds %d elements, the call needs %d" % (name, a.size, count))


def _render_colors(image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c, triangle_id):
  image = _buffer('image', image, np.uint8, 1)
  face_mask = _buffer('face_mask', face_mask, np.uint8, 1)
  vertices = _buffer('vertices', vertices, np.float32, 1)
  triangles = _buffer('triangles', triangles, np.int32, 1)
  colors = _buffer('colors', colors, np.float32, 1)
  depth_buffer = _buffer('depth_buffer', depth_buffer, np.float32, 1)
  ntri, h, w, c = int(ntri), int(h), int(w), int(c)
  nver = vertices.size // 3
  _need('image', image, h * w * c)
  _need('face_mask', face_mask, h * w)
  _need('depth_buffer', depth_buffer, h * w)
  _need('triangles', triangles, 3 * ntri)
  _need('colors', colors, c * nver)
  if ntri > 0:
    t = triangles[:3 * ntri]
    if t.min() < 0 or t.max() >= nver:
      raise ValueError('triangle index outside the vertex buffer (the reference would read out of bounds)')
  _lib.check(_lib.lib().vp_render_colors_core(_lib.ptr(image), _lib.ptr(face_mask), _lib.ptr(vertices),
                                              _lib.ptr(triangles), _lib.ptr(colors), _lib.ptr(depth_buffer),
                                              None if triangle_id is None else _lib.ptr(triangle_id),
                                              nver, ntri, h, w, c))


def render_colors_core(image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c):
  """mesh_core_cython.pyx:64-78 -> mesh_core.cpp:169-231 (flat depth, flat colour)."""
  _render_colors(image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c, None)


def render_colors_with_triangle_id(image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c):
  """render_colors_core that also returns the winning triangle per pixel (-1 = untouched), which
  the reference computes implicitly but never stores.  Same argument checks as render_colors_core."""
  triangle_id = np.empty(int(h) * int(w), dtype=np.int32)
  _render_colors(image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c, triangle_id)
  return triangle_id


def rasterize_triangles_core(vertices, triangles, depth_buffer, triangle_buffer, barycentric_weight, nver, ntri, h,
                             w):
  """mesh_core_cython.pyx:49-62 -> mesh_core.cpp:108-166 (interpolated depth, triangle id and
  barycentric weights, the 2-pixel border rule)."""
  vertices = _buffer('vertices', vertices, np.float32, 2)
  triangles = _buffer('triangles', triangles, np.int32, 2)
  depth_buffer = _buffer('depth_buffer', depth_buffer, np.float32, 2)
  triangle_buffer = _buffer('triangle_buffer', triangle_buffer, np.int32, 2)
  barycentric_weight = _buffer('barycentric_weight', barycentric_weight, np.float32, 2)
  nver, ntri, h, w = int(nver), int(ntri), int(h), int(w)
  _need('vertices', vertices, 3 * nver)
  _need('triangles', triangles, 3 * ntri)
  _need('depth_buffer', depth_buffer, h * w)
  _need('triangle_buffer', triangle_buffer, h * w)
  _need('barycentric_weight', barycentric_weight, 3 * h * w)
  if ntri > 0:
    t = triangles.reshape(-1)[:3 * ntri]
    if t.min() < 0 or t.max() >= nver:
      raise ValueError('triangle index outside the vertex buffer (the reference would read out of bounds)')
  _lib.check(_lib.lib().vp_rasterize_triangles_core(_lib.ptr(vertices), _lib.ptr(triangles), _lib.ptr(depth_buffer),
                                                    _lib.ptr(triangle_buffer), _lib.ptr(barycentric_weight), nver,
                                                    ntri, h, w))


def render_texture_core(image, vertices, triangles, texture, tex_coords, tex_triangles, depth_buffer, nver, tex_nver,
                        ntri, h, w, c, tex_h, tex_w, tex_c, mapping_type):
  """mesh_core_cython.pyx:80-99 -> mesh_core.cpp:234-333 (z-buffer of rasterize_triangles_core, the
  winner's texel sampled nearest / bilinearly; the texture y coordinate is read with the mesh vertex
  index like the reference does, mesh_core.cpp:270-272)."""
  image = _buffer('image', image, np.float32, 3)
  vertices = _buffer('vertices', vertices, np.float32, 2)
  triangles = _buffer('triangles', triangles, np.int32, 2)
  texture = _buffer('texture', texture, np.float32, 3)
  tex_coords = _buffer('tex_coords', tex_coords, np.float32, 2)
  tex_triangles = _buffer('tex_triangles', tex_triangles, np.int32, 2)
  depth_buffer = _buffer('depth_buffer', depth_buffer, np.float32, 2)
  nver, tex_nver, ntri, h, w, c = int(nver), int(tex_nver), int(ntri), int(h), int(w), int(c)
  tex_h, tex_w, tex_c, mapping_type = int(tex_h), int(tex_w), int(tex_c), int(mapping_type)
  _need('image', image, h * w * c)
  _need('vertices', vertices, 3 * nver)
  _need('triangles', triangles, 3 * ntri)
  _need('texture', texture, tex_h * tex_w * tex_c)
  _need('tex_coords', tex_coords, 3 * tex_nver)
  _need('tex_triangles', tex_triangles, 3 * ntri)
  _need('depth_buffer', depth_buffer, h * w)
  _lib.check(_lib.lib().vp_render_texture_core(_lib.ptr(image), _lib.ptr(vertices), _lib.ptr(triangles),
                                               _lib.ptr(texture), _lib.ptr(tex_coords), _lib.ptr(tex_triangles),
                                               _lib.ptr(depth_buffer), nver, tex_nver, ntri, h, w, c, tex_h, tex_w,
                                               tex_c, mapping_type))


def get_normal_core(normal, tri_normal, triangles, ntri):
  """mesh_core_cython.pyx:40-47 -> mesh_core.cpp:85-105: normal[v] += tri_normal[i] over the corners of every
  triangle, in ascending triangle order (float32, bit-identical to the reference's loop)."""
  normal = _buffer('normal', normal, np.float32, 2)
  tri_normal = _buffer('tri_normal', tri_normal, np.float32, 2)
  triangles = _buffer('triangles', triangles, np.int32, 2)
  ntri = int(ntri)
  _need('tri_normal', tri_normal, 3 * ntri)
  _need('triangles', triangles, 3 * ntri)
  _lib.check(_lib.lib().vp_get_normal_core(_lib.ptr(normal), _lib.ptr(tri_normal), _lib.ptr(triangles),
                                           normal.size // 3, ntri))
