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
print('sanitize step ok', int(a[::2, ::16, ::16].sum()))
