"""ctypes binding of libvpb200.so (the C ABI declared in include/vpb200.h).

There is no CPU fallback: if the shared library is missing, importing a compute entry point
raises; if there is no CUDA device, every compute call returns VP_ERR_CUDA and raises VpError.
"""
import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libvpb200.so')

VP_OK = 0
F64_MEANSHAPE, F64_IDBASE, F64_EXBASE, F64_MEANTEX, F64_TEXBASE = 1, 2, 4, 8, 16


class VpError(RuntimeError):
  pass


class VpFrames(ctypes.Structure):
  _fields_ = [('nframes', ctypes.c_int),
              ('ex', ctypes.c_void_p),
              ('rotation', ctypes.c_void_p),
              ('translation', ctypes.c_void_p),
              ('gamma', ctypes.c_void_p),
              ('rotate_shape_first', ctypes.c_int),
              ('focal', ctypes.c_double),
              ('center', ctypes.c_double)]


class VpReconOut(ctypes.Structure):
  _fields_ = [('face_shape', ctypes.c_void_p),
              ('face_norm', ctypes.c_void_p),
              ('face_color', ctypes.c_void_p),
              ('projection', ctypes.c_void_p),
              ('z_buffer', ctypes.c_void_p),
              ('flip_y', ctypes.c_int),
              ('image_size', ctypes.c_double)]


# numpy dtype of vp_frame_params (include/vpb200.h), 192 bytes
FRAME_PARAMS_DTYPE = np.dtype([('rotation', np.float64, (9,)), ('translation', np.float32, (3,)),
                               ('gamma', np.float32, (27,))])
assert FRAME_PARAMS_DTYPE.itemsize == 192

_vp = ctypes.c_void_p
_i = ctypes.c_int
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/vpb200.h one to one
SIGNATURES = {
    'vp_last_error': (ctypes.c_char_p, []),
    'vp_version': (_i, []),
    'vp_device_count': (_i, []),
    'vp_host_alloc': (_i, [ctypes.POINTER(_vp), _sz]),
    'vp_host_free': (_i, [_vp]),
    'vp_ipc_export': (_i, [_vp, _vp, ctypes.POINTER(ctypes.c_ulonglong)]),
    'vp_ipc_open': (_i, [_vp, _i, ctypes.POINTER(_vp)]),
    'vp_ipc_close': (_i, [_vp]),
    'vp_copy_async': (_i, [_vp, _vp, _sz, _vp]),
    'vp_peer_signal': (_i, [_vp, ctypes.c_uint, _vp]),
    'vp_peer_wait': (_i, [_vp, _i, ctypes.c_uint, _vp]),
    'vp_peer_timeouts': (_i, []),
    'vp_render_colors_core': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i]),
    'vp_rasterize_triangles_core': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    'vp_render_texture_core': (_i, [_vp] * 7 + [_i] * 10),
    'vp_get_normal_core': (_i, [_vp, _vp, _vp, _i, _i]),
    'vp_composite_axis_table': (_i, [_i, _i, _i, _vp]),
    'vp_composite_placement': (_i, [_i, _i, _i, ctypes.c_double, _vp, _vp, _vp, _vp]),
    'vp_composite_dev': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _vp]),
    'vp_render_colors_batch_dev': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'vp_model_create': (_i, [ctypes.POINTER(_vp), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    'vp_model_destroy': (None, [_vp]),
    'vp_model_nver': (_i, [_vp]),
    'vp_model_ntri': (_i, [_vp]),
    'vp_model_ntiles': (_i, [_vp]),
    'vp_topology_build': (_i, [ctypes.POINTER(_vp), _i, _i, _vp, _vp, _vp]),
    'vp_topology_destroy': (None, [_vp]),
    'vp_topology_sizes': (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    'vp_topology_copy': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'vp_topology_copy_owned': (_i, [_vp, _vp, _vp, _vp]),
    'vp_topology_slot_count': (_i, [_vp]),
    'vp_topology_copy_slots': (_i, [_vp, _vp, _vp, _vp]),
    'vp_set_basis_mode': (_i, [_vp, _i]),
    'vp_set_vertex_mode': (_i, [_vp, _i]),
    'vp_set_raster_path': (_i, [_vp, _i]),
    'vp_model_fused_available': (_i, [_vp]),
    'vp_model_fan_tiles': (_i, [_vp]),
    'vp_set_identity': (_i, [_vp, _vp, _vp]),
    'vp_set_base_shape': (_i, [_vp, _vp]),
    'vp_set_texture': (_i, [_vp, _vp]),
    'vp_get_texture': (_i, [_vp, _vp]),
    'vp_get_base_shape': (_i, [_vp, _vp]),
    'vp_reconstruct': (_i, [_vp, ctypes.POINTER(VpFrames), ctypes.POINTER(VpReconOut)]),
    'vp_illumination': (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    'vp_projection': (_i, [_i, _i, _vp, _vp, _vp, ctypes.c_double, ctypes.c_double, _vp, _vp]),
    'vp_render_sequence': (_i, [_vp, ctypes.POINTER(VpFrames), _i, _vp, _vp, _i, _vp]),
    'vp_render_sequence_dev': (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    'vp_render_sequence_dev_notify': (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _i]),
    'vp_render_sequence_dev_chunks': (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    'vp_basis_dev': (_i, [_vp, _vp, _vp, _i, _vp]),
    'vp_loss_mask_create': (_i, [_vp, _vp, ctypes.POINTER(_vp)]),
    'vp_loss_mask_destroy': (None, [_vp]),
    'vp_expression_loss_dev': (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    'vp_model_rows_pad': (_i, [_vp]),
    'vp_debug_basis_trace': (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    'vp_launch_count': (ctypes.c_ulonglong, []),
    'vp_set_profiling': (_i, [_vp, _i]),
    'vp_get_profile': (_i, [_vp, ctypes.c_char_p, _i, _vp, _i]),
    'vp_get_profile_launches': (_i, [_vp, _vp, _i]),
}

_lib = None


def lib():
  """The loaded library; raises if it has not been built (python __graft_entry__.py build)."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise VpError('%s is missing: build it with `make -C voicepuppet_b200/csrc` '
                    '(there is no CPU fallback)' % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(handle, name)
      fn.restype = res
      fn.argtypes = args
    _lib = handle
  return _lib


def check(rc):
  if rc != VP_OK:
    msg = lib().vp_last_error()
    raise VpError('libvpb200 error %d: %s' % (rc, msg.decode('utf-8', 'replace') if msg else '?'))


def ptr(a):
  """Address of a C-contiguous numpy array (or None)."""
  if a is None:
    return None
  assert a.flags.c_contiguous
  return ctypes.c_void_p(a.ctypes.data)


def device_count():
  n = lib().vp_device_count()
  return max(n, 0)


def _free_pinned(address):
  try:
    lib().vp_host_free(ctypes.c_void_p(address))
  except Exception:
    pass


def pinned_empty(shape, dtype):
  """Uninitialised numpy array over page-locked host memory (vp_host_alloc).  The memory is
  released when the last view of the array is garbage collected."""
  shape = tuple(int(x) for x in np.atleast_1d(shape))
  nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
  if nbytes == 0:
    return np.empty(shape, dtype=dtype)
  p = ctypes.c_void_p()
  check(lib().vp_host_alloc(ctypes.byref(p), nbytes))
  buf = (ctypes.c_ubyte * nbytes).from_address(p.value)
  weakref.finalize(buf, _free_pinned, p.value)
  return np.frombuffer(buf, dtype=dtype).reshape(shape)
