"""Timeline of one push-gather step per rank (torchrun, N >= 2): where the step time goes."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from voicepuppet_b200 import _lib, render, synthetic
from voicepuppet_b200.model import DeviceModel

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
frames, res = 75, 256
model = synthetic.cached_model()
dm = DeviceModel.of(model, lr)
coeffs = synthetic.make_coeffs(frames * world, seed=1)[rank * frames:(rank + 1) * frames]
angles = render.jitter_angle_sequence(frames * world)[rank * frames:(rank + 1) * frames]
dm.set_identity(coeffs[0:1, :80], coeffs[0:1, 144:224])
ex_dev, params_dev = render.device_inputs(coeffs, angles, dev)
peer = render.PeerFrameBuffer(frames, res, world, rank, dev)
lib = _lib.lib()
align = torch.zeros(1, dtype=torch.int32, device=dev)
plans = [[75], [25, 25, 25], [9, 29, 28, 9]]
for plan in plans:
  for it in range(6):
    torch.cuda.synchronize(); dist.barrier()
    dist.all_reduce(align)
    compute = torch.cuda.current_stream(dev)
    copy = render._comm_stream(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e_start = ev(); e_start.record()
    marks = []
    peer.step += 1
    if rank == 0:
      render.render_device(dm, ex_dev, params_dev, True, res, int(peer.slice_ptr))
      e = ev(); e.record(); marks.append(('own render done', e))
      _lib.check(lib.vp_peer_signal(ctypes.c_void_p(peer.flags_ptr), peer.step, ctypes.c_void_p(compute.cuda_stream)))
      _lib.check(lib.vp_peer_wait(ctypes.c_void_p(peer.flags_ptr), world, peer.step, ctypes.c_void_p(compute.cuda_stream)))
      e = ev(); e.record(); marks.append(('all flags seen', e))
    else:
      if peer.local is None:
        peer.local = torch.empty((frames, res, res, 3), dtype=torch.uint8, device=dev)
      events = render.render_device(dm, ex_dev, params_dev, True, res, peer.local, plan=plan)
      # timing twins of the chunk events (torch events from render_device have no timing)
      a = 0
      for c, cev in enumerate(events):
        b = a + plan[c]
        copy.wait_event(cev)
        with torch.cuda.stream(copy):
          e = ev(); e.record(); marks.append(('chunk %d rendered (copy stream)' % c, e))
          _lib.check(lib.vp_copy_async(ctypes.c_void_p(peer.slice_ptr + a * peer.frame_bytes),
                                       ctypes.c_void_p(peer.local.data_ptr() + a * peer.frame_bytes),
                                       (b - a) * peer.frame_bytes, ctypes.c_void_p(copy.cuda_stream)))
          e = ev(); e.record(); marks.append(('chunk %d pushed' % c, e))
        a = b
      with torch.cuda.stream(copy):
        _lib.check(lib.vp_peer_signal(ctypes.c_void_p(peer.flags_ptr + 4 * rank), peer.step, ctypes.c_void_p(copy.cuda_stream)))
        e = ev(); e.record(); marks.append(('signalled', e))
      compute.wait_stream(copy)
    torch.cuda.synchronize()
    if it == 5:
      line = 'plan %s rank %d: ' % (plan, rank) + ', '.join('%s %.0f us' % (n, e_start.elapsed_time(e) * 1e3) for n, e in marks)
      gathered = [None] * world
      dist.all_gather_object(gathered, line)
      if rank == 0:
        for g in gathered[:3]:
          print(g)
dist.barrier()
peer.close()
dist.destroy_process_group()
