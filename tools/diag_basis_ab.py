"""Isolated timing of the tcgen05 basis kernel, cold L2 (write + read flush), T = 75 / 96 / 128 (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, synthetic
from voicepuppet_b200.model import DeviceModel
dm = DeviceModel.of(synthetic.cached_model())
lib = _lib.lib()
dev = torch.device('cuda', 0)
rows_pad = lib.vp_model_rows_pad(dm.handle)
flush_w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_r = torch.ones(64 << 20, dtype=torch.float32, device=dev)
_lib.check(lib.vp_set_basis_mode(dm.handle, 2))
for t in (75, 96, 128):
  ex = torch.randn(t, 64, device=dev)
  disp = torch.empty(t, rows_pad, device=dev)
  bytes_alg = 27424512 + t * (256 + 428508)
  for cond in ('warm', 'cold'):
    ms = []
    for i in range(14):
      if cond == 'cold':
        flush_w.zero_(); flush_r.sum()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      st = torch.cuda.current_stream(dev).cuda_stream
      a.record()
      _lib.check(lib.vp_basis_dev(dm.handle, ex.data_ptr(), disp.data_ptr(), t, st))
      b.record()
      torch.cuda.synchronize()
      if i >= 2:
        ms.append(a.elapsed_time(b))
    m = float(np.median(ms))
    print('EPI=%s T=%3d %-5s %.2f us  %.0f GB/s (%.1f%% of 6548.5)' % (os.environ.get('VPB200_BASIS_EPI', '0'), t, cond, m * 1e3, bytes_alg / m / 1e6, bytes_alg / m / 1e6 / 65.485))
