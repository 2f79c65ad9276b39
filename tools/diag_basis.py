"""Diagnostic: accuracy pattern and isolated timing of the basis kernels (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, synthetic
from voicepuppet_b200.model import DeviceModel

model = synthetic.cached_model()
dm = DeviceModel.of(model)
lib = _lib.lib()
for t in (75, 16, 128):
  coeffs = synthetic.make_coeffs(t, seed=3)
  dm.set_identity(coeffs[0:1, :80], coeffs[0:1, 144:224])
  base = dm.get_base_shape()
  eye = np.tile(np.eye(3).reshape(1, 9), (t, 1))
  z3, z27 = np.zeros((t, 3), np.float32), np.zeros((t, 27), np.float32)
  want = base[None] + np.einsum('ij,tj->ti', model.exBase.astype(np.float32).astype(np.float64),
                                coeffs[:, 80:144].astype(np.float64)).reshape(t, -1, 3)
  for mode in (1, 2):
    _lib.check(lib.vp_set_basis_mode(dm.handle, mode))
    got = dm.reconstruct(coeffs[:, 80:144], eye, z3, z27, want=('shape',))['shape']
    err = np.abs(got - want)
    print('T=%d mode=%d max err %.3e mean err %.3e' % (t, mode, err.max(), err.mean()))
    if err.max() > 1e-6:
      bad = np.argwhere(err > 1e-6)
      print('  bad count', len(bad), 'frames', np.unique(bad[:, 0])[:20], 'verts (orig) first', np.unique(bad[:, 1])[:10])
      per_frame = err.reshape(t, -1).max(axis=1)
      print('  per-frame max', np.round(per_frame[:20], 6))
_lib.check(lib.vp_set_basis_mode(dm.handle, 0))
