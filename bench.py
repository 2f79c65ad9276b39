#!/usr/bin/env python
"""Benchmark of the BFM reconstruction + rasterization hot path (BASELINE.json: rendered frames/s).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one JSON line)
  python bench.py --impl reference ...                          the reference's CPU path, all host cores

A "step" is one pass of the hot path over one batch of synthetic coefficients: the GRID-utterance
configuration (75 frames at 256x256, BASELINE.json configs[1]) per GPU.  With N > 1 every rank
renders its own 75-frame shard (weak scaling, no data-path collective) and the frames are gathered
to rank 0 over NCCL inside the timed step.

  value     frames/s, inputs (expression coefficients, per-frame parameters) resident in HBM,
            outputs left in HBM; timed with CUDA events per step, L2 flushed between steps
  e2e       frames/s through voicepuppet_b200.render.render_sequence with host coefficient rows in,
            rendered frames out in page-locked host memory (h2d + kernels + d2h inside the timing)
  roofline  dominant kernel: algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the reference algorithm (oracle numpy reconstruction + the reference's own C++
            rasterizer when it was compiled, else its C restatement) on all host cores, same run
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VER, N_TRI = 35709, 70789
METRIC = 'rendered frames/sec (BFM recon+raster)'


# ---------------------------------------------------------------------------------------------
# algorithmic bytes (restated in DESIGN.md)
# ---------------------------------------------------------------------------------------------
def algorithmic_bytes(t, res):
  v = 3 * N_VER * 4                      # one float32 xyz (or rgb) array per frame: 428,508 B
  px = res * res
  basis = 4 * 64 * 3 * N_VER             # 27,424,512
  tri = 4 * 3 * N_TRI
  ring = 4 * 8 * N_VER
  per_kernel = {
      'basis': basis + t * (256 + v),                               # read exBase once, write shape per frame
      'vertex': v + v + tri + ring + t * (v + v + v + 192),         # id-shape, texture, adjacency; shape in, vertices + colours out
      'scatter': tri + t * (v + 8 * px),                            # vertices in, z-buffer keys initialised/updated
      'resolve': t * (8 * px + v + 4 * px),                         # keys in, colours in, image + mask out
  }
  total = 30273684 + t * (2572076 + 20 * px)
  return per_kernel, total


def ncu_traffic():
  """Per-launch DRAM traffic of the hot kernels from the committed ncu --set full capture of this workload
  (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None when absent or for another workload."""
  try:
    with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
      return json.load(f)
  except Exception:
    return {}


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
  QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index = index
    self.rows = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                    '--format=csv,noheader,nounits', '-lms', '100'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._pump, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.rows.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
    for row in self.rows:
      parts = [p.strip() for p in row.split(',')]
      if len(parts) < 9:
        continue
      try:
        sm.append(float(parts[1]))
        mx.append(float(parts[2]))
      except ValueError:
        continue
      for name, val in zip(names, parts[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# the reference arm / CPU baseline
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(frames, res, steps, warmup):
  """Times the reference's CPU algorithm on all host cores; each step renders `frames` frames."""
  from oracle import pipeline, reconstruct_oracle as orc
  from voicepuppet_b200 import synthetic
  cores = os.cpu_count() or 1
  workers = max(1, min(cores, frames))
  synthetic.cached_model()                                    # build / cache before forking
  coeffs = synthetic.make_coeffs(frames, seed=1)
  angles = orc.jitter_angle_sequence(frames)[:, 0, :]
  raster_kind = pipeline.rasterizer()[1]
  # reconstruct_mesh.py is Python and does not exist on the GPU box: its numpy restatement (pinned bit for bit
  # to the live reference, tests/test_oracle_reconstruct.py) runs instead -> "port"; the rasterizer is the
  # reference's own mesh_core.cpp compiled in place when oracle/_ref travelled with the snapshot
  kind = 'port'
  times = []
  import multiprocessing as mp
  ctx = mp.get_context('fork')
  bounds = np.linspace(0, frames, workers + 1).astype(int)
  jobs = [(coeffs[a:b], angles[a:b]) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
  with ctx.Pool(workers, initializer=pipeline._pool_init, initargs=({}, res)) as pool:
    pool.map(pipeline._pool_work, [(coeffs[:1], angles[:1])] * workers)
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      pool.map(pipeline._pool_work, jobs, chunksize=1)
      dt = time.perf_counter() - t0
      if i >= warmup:
        times.append(dt)
  sec = sum(times) / len(times)
  # "as shipped": the reference's frame loop is one Python thread, one frame per iteration
  n1 = min(frames, 12)
  pipeline._pool_init({}, res)
  pipeline._pool_work((coeffs[:1], angles[:1]))
  t0 = time.perf_counter()
  pipeline._pool_work((coeffs[:n1], angles[:n1]))
  single = n1 / (time.perf_counter() - t0)
  return {'fps': frames / sec, 'sec_per_step': sec, 'cores': workers, 'kind': kind, 'frames': frames,
          'raster_kind': raster_kind, 'single_core_fps': single, 'single_core_frames': n1}


def run_reference_arm(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  r = cpu_reference_run(args.frames, args.res, args.steps, args.warmup)
  sample = '%d frames at %dx%d per step over %d worker processes (%s rasterizer)' % (
      r['frames'], args.res, args.res, r['cores'],
      "numpy restatement of reconstruct_mesh.py + the reference's own mesh_core.cpp" if r['raster_kind'] == 'reference'
      else 'numpy restatement of reconstruct_mesh.py + C restatement of mesh_core.cpp')
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': r['fps'], 'unit': 'frames/s', 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * r['sec_per_step'], 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': 'GRID utterance: %d frames at %dx%d (BASELINE.json configs[1])' % (args.frames, args.res, args.res),
                 'model': 'synthetic BFM-shaped model, 35709 vertices / 70789 triangles, seed 0', 'coeff_seed': 1},
      'cpu_baseline': {'value': r['fps'], 'unit': 'frames/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': sample,
                       'single_core': {'value': r['single_core_fps'], 'unit': 'frames/s',
                                       'sample': '%d frames, one process, one frame per iteration (the frame loop as shipped)' % r['single_core_frames']}},
      'e2e': {'value': r['fps'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
  import torch
  import torch.distributed as dist
  from voicepuppet_b200 import _lib, render, synthetic
  from voicepuppet_b200.model import DeviceModel, rotation_matrices

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  t_local, res = args.frames, args.res
  t_total = t_local * world
  model = synthetic.cached_model()
  dm = DeviceModel.of(model, local_rank)
  coeffs_all = synthetic.make_coeffs(t_total, seed=1)
  angles_all = render.jitter_angle_sequence(t_total)
  begin = rank * t_local
  coeffs = coeffs_all[begin:begin + t_local]
  angles = angles_all[begin:begin + t_local]
  dm.set_identity(coeffs[0:1, :80], coeffs[0:1, 144:224])

  # device-resident inputs
  ex_dev, params_dev = render.device_inputs(coeffs, angles, dev)
  frames_dev = torch.empty((t_local, res, res, 3), dtype=torch.uint8, device=dev)
  mask_dev = torch.empty((t_local, res, res), dtype=torch.uint8, device=dev)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
  flush_src = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MiB read pass
  flush_mode = os.environ.get('VPB200_FLUSH', 'write+read')

  def flush_l2():
    # a 256 MiB write evicts everything; the optional 256 MiB read afterwards pushes the flush's own dirty
    # lines out to HBM, so the timed step does not pay for writing back the flush buffer
    flush.zero_()
    if flush_mode == 'write+read':
      flush_src.sum()
  lib = _lib.lib()
  npix = res * res

  gather_mode = os.environ.get('VPB200_GATHER', 'p2p')
  peer = None
  if world > 1 and gather_mode == 'p2p':
    peer = render.PeerFrameBuffer(t_local, res, world, rank, dev)

  def step():
    if world == 1:
      render.render_device(dm, ex_dev, params_dev, True, res, frames_dev, mask_dev)
    elif peer is not None:   # every rank's resolve kernel stores straight into rank 0's buffer over NVLink
      peer.render_into(dm, ex_dev, params_dev, True, mode=os.environ.get('VPB200_PEER_MODE', 'auto'),
                       notify_frames=int(os.environ['VPB200_NOTIFY_FRAMES']) if 'VPB200_NOTIFY_FRAMES' in os.environ else None)
    else:                    # baseline: render locally, NCCL gather to rank 0
      render.pipelined_gather(dm, ex_dev, params_dev, True, res, frames_dev, world, rank)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  for _ in range(max(args.warmup, 3)):
    flush_l2()
    step()
  barrier()

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  align = torch.zeros(1, dtype=torch.int32, device=dev)
  launches0 = lib.vp_launch_count()
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
  barrier()
  for i in range(args.steps):
    flush_l2()                         # L2 flush between timed steps (outside the per-step events)
    if world > 1:
      dist.all_reduce(align)           # device-side barrier: every rank's timed step starts together
    starts[i].record()
    step()
    ends[i].record()
  barrier()
  launches = lib.vp_launch_count() - launches0
  step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
  total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
  if world > 1:
    dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
  total_ms = float(total_ms.item())
  # keep the GPU busy a little longer so the clock sampler sees the load
  t_end = time.time() + 1.0
  while time.time() < t_end:
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.vp_render_sequence_dev(dm.handle, t_local, ex_dev.data_ptr(), params_dev.data_ptr(), 1, res,
                                          frames_dev.data_ptr(), mask_dev.data_ptr(), stream))
    torch.cuda.synchronize(dev)
  barrier()
  clocks = sampler.stop() if rank == 0 else None

  # per-kernel durations (CUDA events inside the library, same stream), L2 flushed between passes
  prof = {}
  if rank == 0:
    dm.set_profiling(True)
    acc = {}
    n_prof = max(3, min(args.steps, 10))
    for _ in range(n_prof):
      flush_l2()
      stream = torch.cuda.current_stream(dev).cuda_stream
      _lib.check(lib.vp_render_sequence_dev(dm.handle, t_local, ex_dev.data_ptr(), params_dev.data_ptr(), 1, res,
                                            frames_dev.data_ptr(), mask_dev.data_ptr(), stream))
      for k, v in dm.profile().items():
        acc[k] = acc.get(k, 0.0) + v
    dm.set_profiling(False)
    prof = {k: v / n_prof for k, v in acc.items()}

  # end to end through the public API: host coefficients in, frames in page-locked host memory out
  e2e = None
  if True:
    out_host = _lib.pinned_empty((t_local, res, res, 3), np.uint8)
    for _ in range(3):
      render.render_sequence(coeffs, model, res=res, angles=angles, device=local_rank, out=out_host)
    barrier()
    n_e2e = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
      render.render_sequence(coeffs, model, res=res, angles=angles, device=local_rank, out=out_host)
    torch.cuda.synchronize(dev)
    e2e_sec = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(e2e_sec, op=dist.ReduceOp.MAX)
    e2e = t_total / float(e2e_sec.item())
    checksum = int(np.asarray(out_host[::7, ::8, ::8]).sum())
    same = bool(np.array_equal(np.asarray(out_host[0]), frames_dev[0].cpu().numpy()))

  if rank == 0:
    per_kernel_bytes, total_bytes = algorithmic_bytes(t_local, res)
    peak, peak_src = measured_peaks()
    ms_per_step = total_ms / args.steps
    value = t_total / (ms_per_step * 1e-3)
    dominant = max(prof, key=prof.get) if prof else None
    roofline = None
    kernels = {}
    for k, ms in prof.items():
      if ms > 0:
        gbs = per_kernel_bytes[k] / (ms * 1e-3) / 1e9
        kernels[k] = {'ms': round(ms, 5), 'algorithmic_bytes': per_kernel_bytes[k], 'gbs': round(gbs, 1),
                      'frac': round(gbs / peak, 4)}
    if dominant:
      traffic = ncu_traffic() if (t_local == 75 and res == 256) else {}
      for k in kernels:
        kernels[k]['traffic'] = traffic.get(k)
      roofline = {'bound': 'hbm', 'kernel': dominant, 'achieved': kernels[dominant]['gbs'], 'peak': peak,
                  'unit': 'GB/s', 'frac': kernels[dominant]['frac'], 'traffic': traffic.get(dominant),
                  'traffic_source': traffic.get('source'), 'peak_source': peak_src,
                  'note': 'dominant kernel by duration; it is bound by the LSU data pipe / instruction issue, not by '
                          'HBM (profiles/): the HBM-bound kernels of the path are basis and resolve, see "kernels"'}
    pipeline_gbs = total_bytes / (ms_per_step * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
      r = cpu_reference_run(args.cpu_frames, res, 1, 1)
      cpu = {'value': r['fps'], 'unit': 'frames/s', 'cores': r['cores'], 'kind': r['kind'],
             'single_core': {'value': r['single_core_fps'], 'unit': 'frames/s',
                             'sample': '%d frames, one process, one frame per iteration (the frame loop as shipped)' % r['single_core_frames']},
             'sample': '%d frames at %dx%d, one pass over %d worker processes (numpy restatement of reconstruct_mesh.py'
                       ' + %s)' % (r['frames'], res, res, r['cores'],
                                   "the reference's own mesh_core.cpp" if r['raster_kind'] == 'reference'
                                   else 'C restatement of mesh_core.cpp')}
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'GRID utterance: %d frames at %dx%d per GPU (BASELINE.json configs[1])' % (t_local, res, res),
                   'frames_per_gpu': t_local, 'resolution': res,
                   'model': 'synthetic BFM-shaped model, 35709 vertices / 70789 triangles, seed 0', 'coeff_seed': 1,
                   'l2': 'flushed between timed steps (256 MiB write' + (', then 256 MiB read so no dirty lines remain)' if flush_mode == 'write+read' else ')'),
                   'gather': ('none' if world == 1 else ('finished chunks of frames pushed into rank 0 buffer over NVLink (CUDA IPC peer memory, copy engine) under the rendering of the next chunk; device-side completion flags' if peer is not None else 'NCCL gather of uint8 frames to rank 0 inside the step'))},
        'roofline': roofline,
        'roofline_pipeline': {'algorithmic_bytes': total_bytes, 'achieved': round(pipeline_gbs, 1), 'peak': peak,
                              'unit': 'GB/s', 'frac': round(pipeline_gbs / peak, 4)},
        'kernels': kernels,
        'cpu_baseline': cpu,
        'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': t_local * (64 * 4 + 192),
                'd2h_bytes_per_step': t_local * res * res * 3, 'matches_device_run': same, 'checksum': checksum},
        'gpu_launches': int(launches),
        'clocks': clocks,
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--frames', type=int, default=75, help='frames per GPU per step')
  ap.add_argument('--res', type=int, default=256)
  ap.add_argument('--cpu-frames', type=int, default=300, help='frames of the bounded CPU-baseline sample')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference_arm(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
