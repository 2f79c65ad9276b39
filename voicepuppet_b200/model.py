"""Device-resident twin of the reference's BFM model object.

``DeviceModel.of(facemodel)`` accepts anything with the 8 attributes of
``utils/bfm_load_data.py:9-21`` (class BFM): meanshape, idBase, exBase, meantex, texBase,
tri (1-based), point_buf (1-based, pad F+1), keypoints; uploads it once and caches the handle
per (object, device).  All arithmetic happens in libvpb200.so.
"""
import ctypes
import weakref

import numpy as np

from . import _lib

N_ID, N_EX, N_TEX = 80, 64, 80
IMG = 224
FOCAL = 1015.0
CENTER = 112.0


def _as_float_array(a):
  a = np.asarray(a)
  if a.dtype not in (np.float32, np.float64):
    a = a.astype(np.float64)
  return np.ascontiguousarray(a)


def rotation_matrices(angles):
  """Compute_rotation_matrix (reference utils/reconstruct_mesh.py:68-91) for [T,3] float32 angles
  -> [T,3,3] float64.  cos/sin are evaluated in the angles' own dtype (float32 in the reference's
  callers) and the products in float64, exactly like the reference's np.array([...]) construction."""
  angles = np.asarray(angles)
  if angles.ndim == 1:
    angles = angles[None, :]
  t = angles.shape[0]
  c = np.cos(angles).astype(np.float64)
  s = np.sin(angles).astype(np.float64)
  rx = np.zeros((t, 3, 3))
  ry = np.zeros((t, 3, 3))
  rz = np.zeros((t, 3, 3))
  rx[:, 0, 0] = 1.0
  rx[:, 1, 1] = c[:, 0]
  rx[:, 1, 2] = -s[:, 0]
  rx[:, 2, 1] = s[:, 0]
  rx[:, 2, 2] = c[:, 0]
  ry[:, 0, 0] = c[:, 1]
  ry[:, 0, 2] = s[:, 1]
  ry[:, 1, 1] = 1.0
  ry[:, 2, 0] = -s[:, 1]
  ry[:, 2, 2] = c[:, 1]
  rz[:, 0, 0] = c[:, 2]
  rz[:, 0, 1] = -s[:, 2]
  rz[:, 1, 0] = s[:, 2]
  rz[:, 1, 1] = c[:, 2]
  rz[:, 2, 2] = 1.0
  rot = np.matmul(np.matmul(rz, ry), rx)
  return np.ascontiguousarray(np.transpose(rot, axes=[0, 2, 1]))


class DeviceModel(object):
  _cache = {}

  def __init__(self, facemodel, device=0):
    lib = _lib.lib()
    meanshape = _as_float_array(facemodel.meanshape).reshape(-1)
    id_base = _as_float_array(facemodel.idBase)
    ex_base = _as_float_array(facemodel.exBase)
    meantex = _as_float_array(facemodel.meantex).reshape(-1)
    tex_base = _as_float_array(facemodel.texBase)
    self.nver = meanshape.shape[0] // 3
    if not (id_base.shape == (3 * self.nver, N_ID) and ex_base.shape == (3 * self.nver, N_EX) and
            tex_base.shape == (3 * self.nver, N_TEX) and meantex.shape == (3 * self.nver,)):
      raise ValueError('facemodel arrays do not have the BFM shapes [3N,80] / [3N,64] / [3N,80] / [1,3N]')
    # (x - 1).astype(np.int32) as in reconstruct_mesh.py:39-40 / infer_bfmvid.py:104
    tri = np.ascontiguousarray((np.asarray(facemodel.tri) - 1).astype(np.int32).reshape(-1, 3))
    point_buf = np.ascontiguousarray((np.asarray(facemodel.point_buf) - 1).astype(np.int32).reshape(self.nver, -1))
    if point_buf.shape[1] != 8:
      raise ValueError('point_buf must be [N,8]')
    self.ntri = tri.shape[0]
    self.keypoints = np.asarray(facemodel.keypoints).astype(np.int64).reshape(-1)
    mask = 0
    for arr, bit in ((meanshape, _lib.F64_MEANSHAPE), (id_base, _lib.F64_IDBASE), (ex_base, _lib.F64_EXBASE),
                     (meantex, _lib.F64_MEANTEX), (tex_base, _lib.F64_TEXBASE)):
      if arr.dtype == np.float64:
        mask |= bit
    # the centre the reference subtracts (reconstruct_mesh.py:27), computed by numpy itself in the
    # mean shape's dtype so that its rounding is reproduced
    center = np.mean(np.reshape(np.asarray(facemodel.meanshape), [1, -1, 3]), axis=1, keepdims=True)
    self.center = np.ascontiguousarray(center.reshape(3).astype(np.float64))
    # dtypes the reference's numpy promotion would give the outputs
    f32 = np.float32
    self.shape_dtype = np.result_type(id_base.dtype, ex_base.dtype, meanshape.dtype, f32)
    self.texture_dtype = np.result_type(tex_base.dtype, meantex.dtype, f32)
    self.device = device
    handle = ctypes.c_void_p()
    _lib.check(lib.vp_model_create(ctypes.byref(handle), device, self.nver, self.ntri, _lib.ptr(meanshape),
                                   _lib.ptr(id_base), _lib.ptr(ex_base), _lib.ptr(meantex), _lib.ptr(tex_base),
                                   mask, _lib.ptr(tri), _lib.ptr(point_buf), _lib.ptr(self.center)))
    self.handle = handle
    self._finalizer = weakref.finalize(self, lib.vp_model_destroy, handle)
    self._identity_key = None

  # -------------------------------------------------------------------------------------
  @classmethod
  def of(cls, facemodel, device=0):
    """Cached DeviceModel for a reference model object (identity-keyed, weakly held)."""
    key = (id(facemodel), device)
    hit = cls._cache.get(key)
    if hit is not None and hit[0]() is facemodel:
      return hit[1]
    dm = cls(facemodel, device)
    try:
      ref = weakref.ref(facemodel, lambda _r, k=key: cls._cache.pop(k, None))
    except TypeError:
      ref = (lambda fm=facemodel: fm)
    cls._cache[key] = (ref, dm)
    return dm

  def close(self):
    self._finalizer()

  # -------------------------------------------------------------------------------------
  def set_identity(self, id_coeff=None, tex_coeff=None):
    """Per-clip constants: base shape = meanshape + idBase.id - centre, texture = meantex + texBase.tex."""
    idc = None if id_coeff is None else np.ascontiguousarray(np.asarray(id_coeff, dtype=np.float32).reshape(N_ID))
    texc = None if tex_coeff is None else np.ascontiguousarray(np.asarray(tex_coeff, dtype=np.float32).reshape(N_TEX))
    key = (None if idc is None else idc.tobytes(), None if texc is None else texc.tobytes())
    if self._identity_key is not None and key == self._identity_key and None not in key:
      return
    _lib.check(_lib.lib().vp_set_identity(self.handle, _lib.ptr(idc), _lib.ptr(texc)))
    self._identity_key = key if None not in key else None

  def set_base_shape(self, shape):
    shape = np.ascontiguousarray(np.asarray(shape, dtype=np.float64).reshape(self.nver, 3))
    _lib.check(_lib.lib().vp_set_base_shape(self.handle, _lib.ptr(shape)))
    self._identity_key = None

  def set_texture(self, texture):
    texture = np.ascontiguousarray(np.asarray(texture, dtype=np.float32).reshape(self.nver, 3))
    _lib.check(_lib.lib().vp_set_texture(self.handle, _lib.ptr(texture)))
    self._identity_key = None

  def get_texture(self):
    out = np.empty((self.nver, 3), dtype=np.float32)
    _lib.check(_lib.lib().vp_get_texture(self.handle, _lib.ptr(out)))
    return out

  def get_base_shape(self):
    out = np.empty((self.nver, 3), dtype=np.float64)
    _lib.check(_lib.lib().vp_get_base_shape(self.handle, _lib.ptr(out)))
    return out

  # -------------------------------------------------------------------------------------
  @staticmethod
  def _frames(ex, rotation, translation, gamma, rotate_shape_first, focal=FOCAL, center=CENTER):
    t = rotation.shape[0]
    keep = []

    def c(a, dtype, width):
      if a is None:
        return None
      a = np.ascontiguousarray(np.asarray(a, dtype=dtype).reshape(t, width))
      keep.append(a)
      return a

    fr = _lib.VpFrames()
    fr.nframes = t
    exa = c(ex, np.float32, N_EX)
    fr.ex = None if exa is None else exa.ctypes.data
    fr.rotation = c(rotation, np.float64, 9).ctypes.data
    fr.translation = c(translation, np.float32, 3).ctypes.data
    fr.gamma = c(gamma, np.float32, 27).ctypes.data
    fr.rotate_shape_first = int(bool(rotate_shape_first))
    fr.focal = float(focal)
    fr.center = float(center)
    return fr, keep

  def reconstruct(self, ex, rotation, translation, gamma, rotate_shape_first=False, want=('shape', 'norm', 'color',
                                                                                         'projection', 'zbuffer'),
                  flip_y=True, focal=FOCAL, center=CENTER, image_size=float(IMG)):
    """Batched Reconstruction / Reconstruction_rotation.  Returns a dict of [T,N,k] arrays."""
    rotation = np.asarray(rotation, dtype=np.float64).reshape(-1, 9)
    t = rotation.shape[0]
    fr, keep = self._frames(ex, rotation, translation, gamma, rotate_shape_first, focal, center)
    out = _lib.VpReconOut()
    res = {}
    spec = (('shape', 'face_shape', np.float64, 3), ('norm', 'face_norm', np.float32, 3),
            ('color', 'face_color', np.float32, 3), ('projection', 'projection', np.float64, 2),
            ('zbuffer', 'z_buffer', np.float64, 1))
    for name, field, dtype, k in spec:
      if name in want:
        res[name] = np.empty((t, self.nver, k), dtype=dtype)
        setattr(out, field, res[name].ctypes.data)
    out.flip_y = int(bool(flip_y))
    out.image_size = float(image_size)
    _lib.check(_lib.lib().vp_reconstruct(self.handle, ctypes.byref(fr), ctypes.byref(out)))
    del keep
    return res

  def render_sequence(self, ex, rotation, translation, gamma, res=IMG, rotate_shape_first=True, want_mask=False,
                      out=None, mask_out=None):
    """The fused path: per-frame inputs -> uint8 frames [T,res,res,3] (and masks [T,res,res]).
    ``out`` / ``mask_out`` may be numpy arrays (host; page-locked ones make the drain asynchronous)
    or torch CUDA tensors on this model's device (frames stay on the GPU, e.g. for the NCCL gather)."""
    rotation = np.asarray(rotation, dtype=np.float64).reshape(-1, 9)
    t = rotation.shape[0]
    fr, keep = self._frames(ex, rotation, translation, gamma, rotate_shape_first)
    if out is None:
      out = _lib.pinned_empty((t, res, res, 3), np.uint8)
    on_device = hasattr(out, 'data_ptr')
    if want_mask and mask_out is None:
      if on_device:
        raise ValueError('pass mask_out explicitly when rendering into device memory')
      mask_out = _lib.pinned_empty((t, res, res), np.uint8)

    def address(a, count):
      if a is None:
        return None
      if on_device:
        if not (a.is_cuda and a.is_contiguous() and a.numel() == count and a.element_size() == 1):
          raise ValueError('device output must be a contiguous uint8 CUDA tensor of %d elements' % count)
        if a.device.index != self.device:
          raise ValueError('device output lives on cuda:%s, the model on cuda:%d' % (a.device.index, self.device))
        return ctypes.c_void_p(a.data_ptr())
      if not (a.dtype == np.uint8 and a.flags.c_contiguous and a.size == count):
        raise ValueError('host output must be a C-contiguous uint8 array of %d elements' % count)
      return _lib.ptr(a)

    stream = None
    if on_device:
      import torch
      stream = ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
    _lib.check(_lib.lib().vp_render_sequence(self.handle, ctypes.byref(fr), int(res), address(out, t * res * res * 3),
                                             address(mask_out, t * res * res), int(on_device), stream))
    del keep
    return (out, mask_out) if want_mask else out

  def set_profiling(self, enabled):
    _lib.check(_lib.lib().vp_set_profiling(self.handle, int(bool(enabled))))

  def profile(self):
    names = ctypes.create_string_buffer(256)
    ms = np.zeros(8, dtype=np.float32)
    _lib.check(_lib.lib().vp_get_profile(self.handle, names, 256, _lib.ptr(ms), 8))
    keys = names.value.decode().split(';')
    return dict(zip(keys, [float(x) for x in ms[:len(keys)]]))

  def profile_launches(self):
    """Kernel launches behind every entry of profile() (same keys)."""
    names = ctypes.create_string_buffer(256)
    ms = np.zeros(8, dtype=np.float32)
    _lib.check(_lib.lib().vp_get_profile(self.handle, names, 256, _lib.ptr(ms), 8))
    counts = np.zeros(8, dtype=np.int32)
    _lib.check(_lib.lib().vp_get_profile_launches(self.handle, _lib.ptr(counts), 8))
    keys = names.value.decode().split(';')
    return dict(zip(keys, [int(x) for x in counts[:len(keys)]]))
