"""How does the tcgen05 basis kernel's time split into fixed and per-tile parts? (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, synthetic
from voicepuppet_b200.model import DeviceModel

lib = _lib.lib()
dev = torch.device('cuda', 0)

def timeit(fn, n=12):
  ms = []
  for i in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    if i >= 2: ms.append(a.elapsed_time(b))
  return float(np.median(ms)) * 1e3

# reference: plain device copies of comparable size
for mb in (8, 30, 60, 120):
  src = torch.empty(mb << 20, dtype=torch.uint8, device=dev); dst = torch.empty_like(src)
  us = timeit(lambda: dst.copy_(src))
  print('torch copy %3d MB -> %3d MB: %.2f us, %.0f GB/s (read+write)' % (mb, mb, us, 2 * (mb << 20) / us / 1e3))
src = torch.empty(60 << 20, dtype=torch.uint8, device=dev)
us = timeit(lambda: src.zero_()); print('torch fill 60 MB: %.2f us, %.0f GB/s' % (us, (60 << 20) / us / 1e3))
e = torch.empty(4, device=dev)
us = timeit(lambda: e.zero_()); print('tiny kernel: %.2f us' % us)

class Bare(object):
  def __init__(self, n):
    rng = np.random.default_rng(0)
    self.meanshape = rng.random((1, 3 * n)).astype(np.float32)
    self.meantex = self.meanshape
    self.idBase = np.zeros((3 * n, 80), np.float32)
    self.texBase = self.idBase
    self.exBase = rng.standard_normal((3 * n, 64)).astype(np.float32)
    self.tri = np.zeros((0, 3)); self.point_buf = np.ones((n, 8)); self.keypoints = np.zeros(68, np.int32)

for n, label in ((6314, '1 tile/CTA'), (12629, '2 tiles/CTA'), (25258, '4 tiles/CTA'), (35709, '5.65'), (50517, '8 tiles/CTA')):
  dm = DeviceModel(Bare(n))
  rows_pad = lib.vp_model_rows_pad(dm.handle)
  for t in (16, 75, 128):
    ex = torch.randn(t, 64, device=dev); disp = torch.empty(t, rows_pad, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    for mode in (2, 1):
      _lib.check(lib.vp_set_basis_mode(dm.handle, mode))
      us = timeit(lambda: _lib.check(lib.vp_basis_dev(dm.handle, ex.data_ptr(), disp.data_ptr(), t, st)))
      mbytes = (rows_pad * 256 + t * (256 + rows_pad * 4)) / 1e6
      print('%-12s tiles=%4d T=%3d mode=%d: %7.2f us  %6.1f MB  %5.0f GB/s' % (label, rows_pad // 128, t, mode, us, mbytes, mbytes / us * 1e3))
  dm.close()
