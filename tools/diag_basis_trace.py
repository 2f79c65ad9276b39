import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from voicepuppet_b200 import _lib, synthetic
from voicepuppet_b200.model import DeviceModel
lib = _lib.lib(); dev = torch.device('cuda', 0)
model = synthetic.cached_model(); dm = DeviceModel.of(model)
rows_pad = lib.vp_model_rows_pad(dm.handle)
st = torch.cuda.current_stream(dev).cuda_stream
for t in (75, 128):
  ex = torch.randn(t, 64, device=dev); disp = torch.empty(t, rows_pad, device=dev)
  trace = torch.zeros(1024, dtype=torch.int64, device=dev)
  for _ in range(3):
    _lib.check(lib.vp_debug_basis_trace(dm.handle, ex.data_ptr(), disp.data_ptr(), t, trace.data_ptr(), st))
  torch.cuda.synchronize()
  full = trace.cpu().numpy()
  tr = full[:256].reshape(4, 16, 4)
  ent, ext = full[256:256 + 296:2], full[257:257 + 296:2]
  e0 = ent[ent > 0].min()
  print('T=%d per-CTA globaltimer (ns after the first CTA entry): entry min/median/max %d / %d / %d, exit min/median/max %d / %d / %d, CTA 0: %d -> %d' % (t, (ent - e0).min(), np.median(ent - e0), (ent - e0).max(), (ext - e0).min(), np.median(ext - e0), (ext - e0).max(), ent[0] - e0, ext[0] - e0))
  t0 = tr[tr > 0].min()
  print('T=%d (clock cycles relative to the first mark)' % t)
  names = ['producer: (it=0: kernel entry, thread-0 prologue done, set-up barrier passed, all roles done) loop-top, afree-ok, issued', 'mma: loop-top, split-ok, accfree-ok, committed',
           'worker: loop-top, full-ok, split-done, arrived', 'epilogue: start, mma-ok, done']
  for r in range(4):
    print(' ', names[r])
    for it in range(7):
      row = tr[r, it]
      if row.max() > 0:
        print('    it=%d ' % it + ' '.join('%7d' % (x - t0) if x > 0 else '      -' for x in row))
