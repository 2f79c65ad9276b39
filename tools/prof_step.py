"""A few device-resident steps of the bench workload (75 frames at 256x256) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voicepuppet_b200 import _lib, render, synthetic
from voicepuppet_b200.model import DeviceModel

frames = int(os.environ.get('FRAMES', '75'))
res = int(os.environ.get('RES', '256'))
steps = int(os.environ.get('STEPS', '4'))
dev = torch.device('cuda', 0)
model = synthetic.cached_model()
dm = DeviceModel.of(model, 0)
coeffs = synthetic.make_coeffs(frames, seed=1)
angles = render.jitter_angle_sequence(frames)
dm.set_identity(coeffs[0:1, :80], coeffs[0:1, 144:224])
ex_dev, params_dev = render.device_inputs(coeffs, angles, dev)
img = torch.empty((frames, res, res, 3), dtype=torch.uint8, device=dev)
msk = torch.empty((frames, res, res), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(steps):
  flush.zero_()
  render.render_device(dm, ex_dev, params_dev, True, res, img, msk)
torch.cuda.synchronize()
print('ok', int(img[::5, ::8, ::8].sum().item()))
