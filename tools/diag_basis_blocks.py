"""K1 (tcgen05 basis kernel) per launch against the frame count: one launch walks ceil(T / 128) frame blocks.
Cold L2 (a 256 MiB write + read between launches), CUDA events around the launch.  (GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, synthetic
from voicepuppet_b200.model import DeviceModel

lib = _lib.lib()
dev = torch.device('cuda', 0)
fm = synthetic.cached_model()
dm = DeviceModel.of(fm)
rows_pad = lib.vp_model_rows_pad(dm.handle)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6548.5
_lib.check(lib.vp_set_basis_mode(dm.handle, 2))
st = torch.cuda.current_stream(dev).cuda_stream
for t in (75, 128, 256, 384, 512, 768, 1024, 2048):
  ex = torch.randn(t, 64, device=dev); disp = torch.empty(t, rows_pad, device=dev)
  us = []
  for i in range(10):
    flush.zero_(); s = int(flush[::4096].sum().item())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _lib.check(lib.vp_basis_dev(dm.handle, ex.data_ptr(), disp.data_ptr(), t, st)); b.record()
    torch.cuda.synchronize()
    if i >= 3: us.append(a.elapsed_time(b) * 1e3)
  us = float(np.median(us))
  mb = (rows_pad * 256 + t * (256 + 428508)) / 1e6
  print('T=%4d  %8.2f us  %7.1f MB algorithmic  %6.0f GB/s  frac %.3f  (%.2f us per 128 frames)' % (t, us, mb, mb / us * 1e3, mb / us * 1e3 / peak, us / t * 128))
_lib.check(lib.vp_set_basis_mode(dm.handle, 0))
