"""Isolated timing of the basis kernels under different L2 pre-conditions (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, synthetic
from voicepuppet_b200.model import DeviceModel

model = synthetic.cached_model()
dm = DeviceModel.of(model)
lib = _lib.lib()
dev = torch.device('cuda', 0)
rows_pad = lib.vp_model_rows_pad(dm.handle)
flush_w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_r = torch.ones(64 << 20, dtype=torch.float32, device=dev)
for t in (75, 128, 16, 8):
  ex = torch.randn(t, 64, device=dev)
  disp = torch.empty(t, rows_pad, device=dev)
  bytes_alg = 27424512 + t * (256 + 428508)
  for mode in (1, 2):
    if t < 16 and mode == 2:
      continue
    _lib.check(lib.vp_set_basis_mode(dm.handle, mode))
    for cond in ('warm', 'write-flush', 'write+read-flush', 'read-flush'):
      ms = []
      for i in range(12):
        if cond in ('write-flush', 'write+read-flush'):
          flush_w.zero_()
        if cond in ('write+read-flush', 'read-flush'):
          flush_r.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream(dev).cuda_stream
        a.record()
        _lib.check(lib.vp_basis_dev(dm.handle, ex.data_ptr(), disp.data_ptr(), t, st))
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
          ms.append(a.elapsed_time(b))
      m = float(np.median(ms))
      print('T=%3d mode=%d %-17s %.2f us  %.0f GB/s (%.0f%% of 6548)' % (t, mode, cond, m * 1e3, bytes_alg / m / 1e6, bytes_alg / m / 1e6 / 65.485))
_lib.check(lib.vp_set_basis_mode(dm.handle, 0))
