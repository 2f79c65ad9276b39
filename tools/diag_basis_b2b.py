"""Back-to-back launches between one event pair: per-launch device time without host launch gaps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, synthetic
from voicepuppet_b200.model import DeviceModel
lib = _lib.lib(); dev = torch.device('cuda', 0)
def b2b(fn, reps=40):
  best = 1e9
  for _ in range(4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big.zero_()            # keeps the GPU busy while the host queues the launches
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / reps * 1e3)
  return best
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
e = torch.empty(4, device=dev)
print('tiny kernel b2b: %.2f us' % b2b(lambda: e.zero_()))
src = torch.empty(30 << 20, dtype=torch.uint8, device=dev); dst = torch.empty_like(src)
print('torch copy 30->30 MB b2b: %.2f us' % b2b(lambda: dst.copy_(src)))
model = synthetic.cached_model(); dm = DeviceModel.of(model)
rows_pad = lib.vp_model_rows_pad(dm.handle)
st = torch.cuda.current_stream(dev).cuda_stream
for t in (16, 75, 128):
  ex = torch.randn(t, 64, device=dev); disp = torch.empty(t, rows_pad, device=dev)
  for mode in (2, 1):
    _lib.check(lib.vp_set_basis_mode(dm.handle, mode))
    us = b2b(lambda: _lib.check(lib.vp_basis_dev(dm.handle, ex.data_ptr(), disp.data_ptr(), t, st)))
    mb = (27424512 + t * (256 + 428508)) / 1e6
    print('T=%3d mode=%d: %.2f us/launch  %.0f GB/s (%.0f%% of 6548.5)' % (t, mode, us, mb / us * 1e3, mb / us * 1e3 / 65.485))
_lib.check(lib.vp_set_basis_mode(dm.handle, 0))
