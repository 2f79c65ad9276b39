"""Drop-in for the reference's ``utils/reconstruct_mesh.py``: same names, same positional
arguments, same return shapes and dtypes -- computed by the CUDA kernels of libvpb200.so.

    from voicepuppet_b200.reconstruct_mesh import *          # instead of utils/reconstruct_mesh

``facemodel`` is the reference's ``BFM`` object (utils/bfm_load_data.py:9-21) or anything with
the same 8 attributes; it is uploaded to the GPU on first use and cached.  Inputs follow the
reference's batch-of-one convention (``coeff`` is ``[1,257]``); the batched forms used by the
fused pipeline live in ``voicepuppet_b200.render``.

Accuracy contract (BASELINE.json north_star): vertices, normals and colours within 1e-5
relative of the reference's numpy result.  The geometry runs in float64 on the device, the
expression contraction and the lighting in float32.
"""
import numpy as np

from . import _lib
from .model import DeviceModel, rotation_matrices

__all__ = ['Split_coeff', 'Shape_formation', 'Compute_norm', 'Texture_formation', 'Compute_rotation_matrix',
           'Projection_layer', 'Illumination_layer', 'Reconstruction', 'Reconstruction_rotation']

_IDENTITY_ROT = np.eye(3, dtype=np.float64).reshape(1, 9)
_ZERO3 = np.zeros((1, 3), dtype=np.float32)
_ZERO27 = np.zeros((1, 27), dtype=np.float32)


def Split_coeff(coeff):
  """reference utils/reconstruct_mesh.py:5-13 (pure slicing; views, like the reference)."""
  id_coeff = coeff[:, :80]
  ex_coeff = coeff[:, 80:144]
  tex_coeff = coeff[:, 144:224]
  angles = coeff[:, 224:227]
  gamma = coeff[:, 227:254]
  translation = coeff[:, 254:]
  return id_coeff, ex_coeff, tex_coeff, angles, gamma, translation


def _first_row(a, width):
  a = np.asarray(a)
  if a.ndim != 2 or a.shape[1] != width:
    raise ValueError('expected an array of shape [1,%d], got %s' % (width, (a.shape,)))
  return a[0:1]


def Shape_formation(id_coeff, ex_coeff, facemodel):
  """reference :20-29 -> [1,N,3]."""
  dm = DeviceModel.of(facemodel)
  dm.set_identity(id_coeff=_first_row(id_coeff, 80))
  out = dm.reconstruct(_first_row(ex_coeff, 64), _IDENTITY_ROT, _ZERO3, _ZERO27, want=('shape',))
  return out['shape'].astype(dm.shape_dtype, copy=False)


def Texture_formation(tex_coeff, facemodel):
  """reference :58-62 -> [1,N,3]."""
  dm = DeviceModel.of(facemodel)
  dm.set_identity(tex_coeff=_first_row(tex_coeff, 80))
  return dm.get_texture().reshape(1, -1, 3).astype(dm.texture_dtype, copy=False)


def Compute_norm(face_shape, facemodel):
  """reference :35-52 -> [1,N,3] float64 unit normals."""
  dm = DeviceModel.of(facemodel)
  face_shape = np.asarray(face_shape)
  dm.set_base_shape(face_shape.reshape(-1, 3))
  out = dm.reconstruct(None, _IDENTITY_ROT, _ZERO3, _ZERO27, want=('norm',))
  return out['norm'].astype(np.float64)


def Compute_rotation_matrix(angles):
  """reference :68-91 -> [1,3,3] float64 (row 0 of ``angles`` only, like the reference)."""
  return rotation_matrices(np.asarray(angles)[0:1])


def Projection_layer(face_shape, rotation, translation, focal=1015.0, center=112.0):
  """reference :100-120 -> (face_projection [1,N,2], z_buffer [1,N,1]) float64."""
  shape = np.ascontiguousarray(np.asarray(face_shape, dtype=np.float64).reshape(-1, 3))
  rot = np.ascontiguousarray(np.asarray(rotation, dtype=np.float64).reshape(-1)[:9])
  trans = np.ascontiguousarray(np.asarray(translation, dtype=np.float32).reshape(-1)[:3])
  n = shape.shape[0]
  proj = np.empty((n, 2), dtype=np.float64)
  zbuf = np.empty((n, 1), dtype=np.float64)
  _lib.check(_lib.lib().vp_projection(0, n, _lib.ptr(shape), _lib.ptr(rot), _lib.ptr(trans), float(focal),
                                      float(center), _lib.ptr(proj), _lib.ptr(zbuf)))
  return proj.reshape(1, n, 2), zbuf.reshape(1, n, 1)


def Illumination_layer(face_texture, norm, gamma):
  """reference :129-168 -> (face_color [1,N,3], lighting [1,N,3]) float64."""
  tex = np.ascontiguousarray(np.asarray(face_texture, dtype=np.float64).reshape(-1, 3))
  nrm = np.ascontiguousarray(np.asarray(norm, dtype=np.float64).reshape(-1, 3))
  g = np.ascontiguousarray(np.asarray(gamma, dtype=np.float32).reshape(-1)[:27])
  n = tex.shape[0]
  color = np.empty((n, 3), dtype=np.float64)
  lighting = np.empty((n, 3), dtype=np.float64)
  _lib.check(_lib.lib().vp_illumination(0, n, _lib.ptr(tex), _lib.ptr(nrm), _lib.ptr(g), _lib.ptr(color),
                                        _lib.ptr(lighting)))
  return color.reshape(1, n, 3), lighting.reshape(1, n, 3)


def _reconstruct(coeff, facemodel, angles, rotate_first):
  dm = DeviceModel.of(facemodel)
  coeff = np.asarray(coeff)
  id_coeff, ex_coeff, tex_coeff, coeff_angles, gamma, translation = Split_coeff(coeff)
  dm.set_identity(_first_row(id_coeff, 80), _first_row(tex_coeff, 80))
  rotation = Compute_rotation_matrix(coeff_angles if angles is None else angles)
  out = dm.reconstruct(ex_coeff[0:1], rotation.reshape(1, 9), translation[0:1], gamma[0:1],
                       rotate_shape_first=rotate_first, want=('shape', 'color', 'projection', 'zbuffer'))
  face_shape = out['shape'].astype(np.float64 if rotate_first else dm.shape_dtype, copy=False)
  face_texture = dm.get_texture().reshape(1, -1, 3).astype(dm.texture_dtype, copy=False)
  face_color = out['color'].astype(np.float64)
  face_projection = out['projection']
  z_buffer = out['zbuffer']
  landmarks_2d = face_projection[:, dm.keypoints, :]
  return face_shape, face_texture, face_color, face_projection, z_buffer, landmarks_2d, translation


def Reconstruction(coeff, facemodel):
  """reference :172-194 -> (face_shape, face_texture, face_color, face_projection, z_buffer,
  landmarks_2d, translation)."""
  return _reconstruct(coeff, facemodel, None, False)


def Reconstruction_rotation(coeff, facemodel, angles):
  """reference :198-223 -> 6-tuple.  The reference's quirks are kept: the coefficient's own angles
  are ignored, the normals are rotated once, the shape is rotated and then rotated again inside
  the projection."""
  return _reconstruct(coeff, facemodel, angles, True)[:6]
