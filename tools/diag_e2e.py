"""Where the end-to-end time goes (run on the GPU box): D2H bandwidth, per-call time vs chunking."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, render, synthetic

dev = torch.device('cuda', 0)
frames, res = 75, 256
model = synthetic.cached_model()
coeffs = synthetic.make_coeffs(frames, seed=1)
angles = render.jitter_angle_sequence(frames)
out = _lib.pinned_empty((frames, res, res, 3), np.uint8)
nbytes = frames * res * res * 3
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
dst = torch.from_numpy(np.asarray(out).reshape(-1))
for _ in range(3):
  dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
  dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print('D2H %d bytes: %.1f us  %.1f GB/s' % (nbytes, dt * 1e6, nbytes / dt / 1e9))
for cf in ('0', '75', '38', '25', '19', '13', '10', '8', '5'):
  if cf == '0':
    os.environ.pop('VPB200_CHUNK_FRAMES', None)
  else:
    os.environ['VPB200_CHUNK_FRAMES'] = cf
  for _ in range(3):
    render.render_sequence(coeffs, model, res=res, angles=angles, device=0, out=out)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  n = 30
  for _ in range(n):
    render.render_sequence(coeffs, model, res=res, angles=angles, device=0, out=out)
  torch.cuda.synchronize()
  dt = (time.perf_counter() - t0) / n
  print('chunk_frames=%s: %.1f us per call, %.0f frames/s' % (cf, dt * 1e6, frames / dt))
# host-side share: the same call with a device output (no D2H)
os.environ.pop('VPB200_CHUNK_FRAMES', None)
dout = torch.empty((frames, res, res, 3), dtype=torch.uint8, device=dev)
for _ in range(3):
  render.render_sequence(coeffs, model, res=res, angles=angles, device=0, out=dout)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
  render.render_sequence(coeffs, model, res=res, angles=angles, device=0, out=dout)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 30
print('device output: %.1f us per call' % (dt * 1e6))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
  render.render_sequence(coeffs, model, res=res, angles=angles, device=0, out=out)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
