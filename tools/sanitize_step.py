"""Small workload for compute-sanitizer (round 2): the 420-vertex model at 512x512 (boxes of hundreds of pixels: the
flattened (triangle,row) walk with its per-warp shared-memory staging), separate kernels and the fused kernel, several
frames per CTA; frames must agree byte for byte."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from voicepuppet_b200 import _lib, render, synthetic
from voicepuppet_b200.model import DeviceModel
model = synthetic.make_model(420, 48)
dm = DeviceModel.of(model)
coeffs = synthetic.make_coeffs(9, seed=3)
a = np.asarray(render.render_sequence(coeffs, model, res=512)).copy()
_lib.check(_lib.lib().vp_set_raster_path(dm.handle, 2))
b = np.asarray(render.render_sequence(coeffs, model, res=512)).copy()
_lib.check(_lib.lib().vp_set_raster_path(dm.handle, 0))
c = np.asarray(render.render_sequence(coeffs[:3], model, res=64)).copy()
assert a.any() and c.any() and np.array_equal(a, b)
# round 2, second half: the GROUP walk of the scatter kernel (from 768x768; a 6000-vertex model at 1024x1024 has boxes of
# 20..100 pixels: group-walked and flattened boxes in the same warps) and the tcgen05 basis kernel walking several frame
# blocks in one launch (300 frames: two full blocks and a partial one)
import torch
model2 = synthetic.make_model(6000, 200)
dm2 = DeviceModel.of(model2)
coeffs2 = synthetic.make_coeffs(5, seed=4)
d = np.asarray(render.render_sequence(coeffs2, model2, res=1024)).copy()
_lib.check(_lib.lib().vp_set_raster_path(dm2.handle, 2))
e = np.asarray(render.render_sequence(coeffs2, model2, res=1024)).copy()
_lib.check(_lib.lib().vp_set_raster_path(dm2.handle, 0))
assert d.any() and np.array_equal(d, e)
lib = _lib.lib()
rows_pad = lib.vp_model_rows_pad(dm2.handle)
dev = torch.device('cuda', 0)
ex = torch.randn(300, 64, device=dev)
out = {}
for mode in (1, 2):
  _lib.check(lib.vp_set_basis_mode(dm2.handle, mode))
  disp = torch.zeros(300, rows_pad, device=dev)
  _lib.check(lib.vp_basis_dev(dm2.handle, ex.data_ptr(), disp.data_ptr(), 300, torch.cuda.current_stream(dev).cuda_stream))
  torch.cuda.synchronize()
  out[mode] = disp.cpu().numpy()
_lib.check(lib.vp_set_basis_mode(dm2.handle, 0))
assert np.abs(out[1] - out[2]).max() < 1e-3
print('sanitize step ok', int(a[::2, ::16, ::16].sum()), int(d[::2, ::16, ::16].sum()))
