#!/usr/bin/env python
"""Benchmark of the BFM reconstruction + rasterization hot path (BASELINE.json: rendered frames/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME]     our CUDA path (one JSON line)
  python bench.py --impl reference ...                                   the reference's CPU path, all host cores

Configurations (BASELINE.json `configs`, in order):
  single        1 frame at 256x256                    configs[0]  (latency of one render_face-sized call)
  grid          75 frames at 256x256                  configs[1]  (one GRID utterance)
  clip1500      1500 frames at 512x512                configs[2]
  sharded12000  12000 frames at 256x256               configs[3]  (frame-sharded over the ranks, gathered to rank 0)
  stress4096    4096 frames at 1024x1024              configs[4]  (the largest single-GPU configuration)
With no --config:  N = 1 runs stress4096 (the largest configuration that fits one GPU) and reports the other four
under "all_configs";  N > 1 runs sharded12000 -- STRONG scaling: the 12000 frames are cut into contiguous shards, every
rank renders its own (no data-path collective) and the frames land in rank 0's buffer over NVLink inside the timed step.

  value     frames/s, inputs (expression coefficients, per-frame parameters) resident in HBM, outputs left in HBM
            (at N > 1: in rank 0's HBM); CUDA events per step, max over ranks, L2 flushed between steps
  e2e       frames/s through the public API with HOST buffers: coefficient rows in, frames out in page-locked host
            memory (h2d + kernels + d2h inside the timing); at N > 1 the step also drains rank 0's gathered buffer
  roofline  dominant kernel: algorithmic bytes per launch / its average CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the reference algorithm (numpy restatement of reconstruct_mesh.py + the reference's own C++ rasterizer
            when it was compiled, else its C restatement) on all host cores, bounded sample of the same workload
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
MODEL_NOTE = 'synthetic BFM-shaped model, 35709 vertices / 70789 triangles, seed 0'

CONFIGS = {   # name -> (frames, resolution, BASELINE.json configs index, description)
    'single': (1, 256, 0, 'single frame at 256x256'),
    'grid': (75, 256, 1, 'GRID utterance: 75 frames at 256x256'),
    'clip1500': (1500, 512, 2, '1-minute clip: 1500 frames at 512x512'),
    'sharded12000': (12000, 256, 3, 'frame-sharded batch: 12000 frames at 256x256'),
    'stress4096': (4096, 1024, 4, 'stress: 4096 frames at 1024x1024 (tcgen05 3xTF32 basis GEMM, tiled raster)'),
}


def workload_name(name):
  _, _, idx, text = CONFIGS[name]
  return '%s (BASELINE.json configs[%d])' % (text, idx) if idx is not None else text


# ---------------------------------------------------------------------------------------------
# algorithmic bytes (restated in DESIGN.md section 6)
# ---------------------------------------------------------------------------------------------
def algorithmic_bytes(t, res, launches=None):
  """Per-kernel and whole-path algorithmic bytes of one step over t frames.  Constants a kernel reads once per
  LAUNCH (the 27 MB expression basis, the per-clip shape / texture, the adjacency) are counted once per launch of
  that kernel (`launches`: kernel -> launches per step, default 1); the whole-path figure counts them once."""
  launches = launches or {}
  v = 3 * N_VER * 4                      # one float32 xyz (or rgb) array per frame: 428,508 B
  px = res * res
  basis = 4 * 64 * 3 * N_VER             # 27,424,512
  tri = 4 * 3 * N_TRI
  ring = 4 * 8 * N_VER
  n = lambda k: max(1.0, float(launches.get(k, 1)))
  per_kernel = {
      'basis': n('basis') * basis + t * (256 + v),                               # exBase per launch, shape out per frame
      'vertex': n('vertex') * (v + v + tri + ring) + t * (v + v + v + 192),      # id-shape, texture, adjacency; shape in, vertices + colours out
      'scatter': n('scatter') * tri + t * (v + 8 * px),                          # vertices in, z-buffer keys initialised / updated
      'resolve': t * (8 * px + v + 4 * px),                                      # keys in, colours in, image + mask out
  }
  # the fused vertex + scatter kernel never materialises the vertex records; it still divides the same figure
  per_kernel['fused'] = n('fused') * (v + v + tri + ring + tri) + t * (v + v + v + 192 + v + 8 * px)
  per_kernel = {k: int(b) for k, b in per_kernel.items()}
  total = 30273684 + t * (2572076 + 20 * px)
  return per_kernel, total


def ncu_traffic(key):
  """Per-launch DRAM traffic of the hot kernels from the committed ncu --set full captures
  (profiles/ncu_traffic.json: {"<frames>x<res>": {kernel: bytes, "source": ...}}); {} when absent."""
  try:
    with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
      d = json.load(f).get(key, {})
      return d if isinstance(d, dict) else {}
  except Exception:
    return {}


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
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
  n1 = min(frames, 8)
  pipeline._pool_init({}, res)
  pipeline._pool_work((coeffs[:1], angles[:1]))
  t0 = time.perf_counter()
  pipeline._pool_work((coeffs[:n1], angles[:n1]))
  single = n1 / (time.perf_counter() - t0)
  return {'fps': frames / sec, 'sec_per_step': sec, 'cores': workers, 'kind': kind, 'frames': frames,
          'raster_kind': raster_kind, 'single_core_fps': single, 'single_core_frames': n1}


def cpu_baseline_dict(r, res, per_step=False):
  sample = '%d frames of the workload at %dx%d %s over %d worker processes (numpy restatement of reconstruct_mesh.py + %s)' % (
      r['frames'], res, res, 'per step' if per_step else 'in one pass', r['cores'],
      "the reference's own mesh_core.cpp" if r['raster_kind'] == 'reference' else 'C restatement of mesh_core.cpp')
  return {'value': r['fps'], 'unit': 'frames/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': sample,
          'single_core': {'value': r['single_core_fps'], 'unit': 'frames/s',
                          'sample': '%d frames, one process, one frame per iteration (the frame loop as shipped)' % r['single_core_frames']}}


def cpu_sample_frames(frames, res, cores, seconds):
  """Frames of a bounded CPU sample worth about `seconds` of wall clock on all the cores: the numpy reconstruction
  costs ~65 ms per frame and core, the C++ rasterizer ~6 ms at 256x256 up to ~50 ms at 1024x1024."""
  per_frame_core_s = 0.07 + 0.05 * (res / 1024.0) ** 2
  n = int(seconds * cores / per_frame_core_s)
  return max(1, min(frames, max(cores, n // cores * cores)))


def default_config(world):
  return 'stress4096' if world == 1 else 'sharded12000'


def run_reference_arm(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  world = max(int(os.environ.get('WORLD_SIZE', '1')), args.gpus)
  name = args.config or default_config(world)
  frames, res = CONFIGS[name][0], CONFIGS[name][1]
  cores = os.cpu_count() or 1
  # every step is a bounded sample of the workload (about 1.5 s on all the cores), so that K + W steps end within minutes
  sample = args.cpu_frames or cpu_sample_frames(frames, res, cores, 1.5)
  r = cpu_reference_run(min(sample, frames), res, args.steps, args.warmup)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': r['fps'], 'unit': 'frames/s', 'n_gpus': args.gpus,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * r['sec_per_step'], 'higher_is_better': True,
      'scaling': 'strong' if name == 'sharded12000' else 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': workload_name(name), 'name': name, 'frames': frames, 'resolution': res,
                 'model': MODEL_NOTE, 'coeff_seed': 1,
                 'sample': 'every step renders a bounded sample of %d of the %d frames on the host cores' % (r['frames'], frames)},
      'cpu_baseline': cpu_baseline_dict(r, res, per_step=True),
      'e2e': {'value': r['fps'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class Env(object):
  """Per-process state shared by the measurements: device, model, L2 flush buffers, distributed group."""

  def __init__(self):
    import torch
    import torch.distributed as dist
    from voicepuppet_b200 import _lib, synthetic
    from voicepuppet_b200.model import DeviceModel
    self.torch, self.dist = torch, dist
    self.world = int(os.environ.get('WORLD_SIZE', '1'))
    self.rank = int(os.environ.get('RANK', '0'))
    self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
      raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(self.local_rank)
    self.dev = torch.device('cuda', self.local_rank)
    if self.world > 1:
      dist.init_process_group('nccl', device_id=self.dev)
    self.model = synthetic.cached_model()
    self.dm = DeviceModel.of(self.model, self.local_rank)
    self.lib = _lib.lib()
    if os.environ.get('VPB200_BENCH_FUSED') == '1':        # A/B: the fused vertex + z-buffer kernel (not the default)
      _lib.check(self.lib.vp_set_raster_path(self.dm.handle, 2))
    self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)       # > 126 MB L2
    self.flush_src = torch.zeros(64 << 20, dtype=torch.float32, device=self.dev)  # 256 MiB read pass
    self.flush_mode = os.environ.get('VPB200_FLUSH', 'write+read')

  def flush_l2(self):
    # a 256 MiB write evicts everything; the 256 MiB read afterwards pushes the flush's own dirty lines out to
    # HBM, so the timed step does not pay for writing back the flush buffer
    self.flush.zero_()
    if self.flush_mode == 'write+read':
      self.flush_src.sum()

  def barrier(self):
    if self.world > 1:
      self.dist.barrier()
    self.torch.cuda.synchronize(self.dev)

  def l2_note(self):
    return 'flushed between timed steps (256 MiB write' + (', then 256 MiB read so no dirty lines remain)'
                                                            if self.flush_mode == 'write+read' else ')')


def measure_single_gpu(env, name, steps, warmup, want_clocks=False, e2e_steps=None):
  """One configuration on this process's GPU: device-resident value, per-kernel profile, end-to-end figure."""
  torch = env.torch
  from voicepuppet_b200 import _lib, render, synthetic
  frames, res = CONFIGS[name][0], CONFIGS[name][1]
  dm, dev, lib = env.dm, env.dev, env.lib
  coeffs = synthetic.make_coeffs(frames, seed=1)
  angles = render.jitter_angle_sequence(frames)
  dm.set_identity(coeffs[0:1, :80], coeffs[0:1, 144:224])
  ex_dev, params_dev = render.device_inputs(coeffs, angles, dev)
  frames_dev = torch.empty((frames, res, res, 3), dtype=torch.uint8, device=dev)
  mask_dev = torch.empty((frames, res, res), dtype=torch.uint8, device=dev)

  def step():
    render.render_device(dm, ex_dev, params_dev, True, res, frames_dev, mask_dev)

  for _ in range(max(warmup, 3)):
    env.flush_l2()
    step()
  torch.cuda.synchronize(dev)
  sampler = ClockSampler(env.local_rank) if want_clocks else None
  if sampler:
    sampler.start()
  launches0 = lib.vp_launch_count()
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
  ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
  torch.cuda.synchronize(dev)
  for i in range(steps):
    env.flush_l2()                         # L2 flush between timed steps (outside the per-step events)
    starts[i].record()
    step()
    ends[i].record()
  torch.cuda.synchronize(dev)
  launches = lib.vp_launch_count() - launches0
  ms_per_step = sum(s.elapsed_time(e) for s, e in zip(starts, ends)) / steps
  clocks = None
  if sampler:
    t_end = time.time() + 1.0              # keep the GPU busy a little longer so the sampler sees the load
    while time.time() < t_end:
      step()
      torch.cuda.synchronize(dev)
    clocks = sampler.stop()

  # per-kernel durations: CUDA events around every launch inside the library, same chunk plan as the timed steps
  # (the only difference: one stream instead of two, so that the spans do not overlap), L2 flushed between passes
  dm.set_profiling(True)
  acc, cnt = {}, {}
  n_prof = max(2, min(steps, 5))
  for _ in range(n_prof):
    env.flush_l2()
    step()
    for k, v in dm.profile().items():
      acc[k] = acc.get(k, 0.0) + v
    for k, v in dm.profile_launches().items():
      cnt[k] = cnt.get(k, 0) + v
  dm.set_profiling(False)
  prof = {k: v / n_prof for k, v in acc.items() if v > 0}
  nlaunch = {k: cnt[k] / n_prof for k in prof}

  # end to end through the public API: host coefficients in, frames in page-locked host memory out
  out_host = _lib.pinned_empty((frames, res, res, 3), np.uint8)
  for _ in range(2):
    render.render_sequence(coeffs, env.model, res=res, angles=angles, device=env.local_rank, out=out_host)
  torch.cuda.synchronize(dev)
  n_e2e = e2e_steps or max(3, min(steps, 10))
  t0 = time.perf_counter()
  for _ in range(n_e2e):
    render.render_sequence(coeffs, env.model, res=res, angles=angles, device=env.local_rank, out=out_host)
  torch.cuda.synchronize(dev)
  e2e_sec = (time.perf_counter() - t0) / n_e2e
  # per clip: a fresh identity every call, so the per-clip contraction (K0: idBase, texBase) is inside the timing
  alt = coeffs.copy()
  t0 = time.perf_counter()
  for i in range(n_e2e):
    alt[:, :80] = coeffs[:, :80] + np.float32(1e-3 * (i + 1))
    render.render_sequence(alt, env.model, res=res, angles=angles, device=env.local_rank, out=out_host)
  torch.cuda.synchronize(dev)
  e2e_clip_sec = (time.perf_counter() - t0) / n_e2e
  render.render_sequence(coeffs, env.model, res=res, angles=angles, device=env.local_rank, out=out_host)
  torch.cuda.synchronize(dev)
  step()
  torch.cuda.synchronize(dev)
  idx = sorted(set([0, frames // 2, frames - 1]))
  same = all(bool(np.array_equal(np.asarray(out_host[i]), frames_dev[i].cpu().numpy())) for i in idx)
  checksum = int(np.asarray(out_host[::max(1, frames // 64), ::8, ::8]).astype(np.int64).sum())
  del out_host, frames_dev, mask_dev
  torch.cuda.empty_cache()

  per_kernel_bytes, total_bytes = algorithmic_bytes(frames, res, nlaunch)
  peak, peak_src = measured_peaks()
  traffic = ncu_traffic('%dx%d' % (frames, res))
  kernels = {}
  for k, ms in prof.items():
    gbs = per_kernel_bytes[k] / (ms * 1e-3) / 1e9
    kernels[k] = {'ms': round(ms, 5), 'launches': round(nlaunch[k], 2), 'ms_per_launch': round(ms / max(nlaunch[k], 1), 5),
                  'algorithmic_bytes': per_kernel_bytes[k], 'gbs': round(gbs, 1), 'frac': round(gbs / peak, 4),
                  'traffic_per_launch': traffic.get(k)}
  pipeline_gbs = total_bytes / (ms_per_step * 1e-3) / 1e9
  out = {
      'workload': workload_name(name), 'frames': frames, 'resolution': res,
      'value': frames / (ms_per_step * 1e-3), 'ms_per_step': ms_per_step,
      'e2e': {'value': frames / e2e_sec, 'unit': 'frames/s', 'h2d_bytes_per_step': frames * (64 * 4 + 192),
              'd2h_bytes_per_step': frames * res * res * 3, 'matches_device_run': same, 'checksum': checksum},
      'e2e_clip': {'value': frames / e2e_clip_sec, 'unit': 'frames/s',
                   'note': 'as e2e, with new identity / texture coefficients every call: the per-clip contraction (K0, '
                           '68 MB of idBase / texBase) and its 640-byte upload are inside the timing'},
      'kernels': kernels,
      'roofline_pipeline': {'algorithmic_bytes': total_bytes, 'achieved': round(pipeline_gbs, 1), 'peak': peak,
                            'unit': 'GB/s', 'frac': round(pipeline_gbs / peak, 4)},
      'gpu_launches': int(launches), 'clocks': clocks,
  }
  if prof:
    dom = max(prof, key=prof.get)
    kd = kernels[dom]
    per_launch = per_kernel_bytes[dom] / max(nlaunch[dom], 1)
    out['roofline'] = {
        'bound': 'hbm', 'kernel': dom, 'achieved': kd['gbs'], 'peak': peak, 'unit': 'GB/s', 'frac': kd['frac'],
        'traffic': traffic.get(dom), 'traffic_source': traffic.get('source'), 'peak_source': peak_src,
        'algorithmic_bytes_per_launch': int(per_launch), 'launch_us': round(1e3 * kd['ms_per_launch'], 2),
        'launches_per_step': kd['launches'],
        'note': 'dominant kernel by duration over the step; durations are CUDA events around every launch of the '
                'timed chunk plan (one stream instead of two so the spans do not overlap); per-kernel figures of the '
                'whole path under "kernels"',
        'limiter': LIMITERS.get(dom)}
  return out


# What ncu and the microbenchmarks say bounds each kernel (profiles/): the HBM fraction above is the
# contract's yardstick, not always the kernel's own limit.
LIMITERS = {
    'basis': 'store path of the displacements (4.4 TB/s of writes per further 128-frame block) + ramp / drain of the launch',
    'vertex': 'instruction issue / LSU (shared-memory position gathers): 68 % issue-active, moves little data',
    'scatter': 'sectors touched by the 64-bit REDG.MAX reductions (1.64 cycles per lane spread, 0.76 in runs of 4: '
               'tools/diag_redg.cu) and then instruction issue (80 % issue-active): moves little data',
    'resolve': 'HBM',
}


def tri_id_report(env):
  """The north_star's end-to-end parity count on a sample of the workload: pixels whose winning triangle differs
  from the CPU reference's, how many of those are depth near-ties (two nearest depths within 1 ulp) and how many
  pixels differ by more than 1/255 in RGB.  CPU oracle as checker, outside every timed region."""
  try:
    from oracle import pipeline, reconstruct_oracle as orc
    from oracle.raster import Oracle
    from voicepuppet_b200 import mesh_core_cython as mc, reconstruct_mesh as rm, render, synthetic
    frames, res = 75, 256
    sample = [0, 37, 74]
    coeffs = synthetic.make_coeffs(frames, seed=1)
    jit = orc.jitter_angle_sequence(frames)
    tris = orc.triangles_flat(env.model)
    got = np.asarray(render.render_sequence(coeffs, env.model, res=res, device=env.local_rank)).copy()
    rep = {'tri_id_mismatch_px': 0, 'near_tie_px': 0, 'edge_flip_px': 0, 'unexplained_px': 0, 'rgb_gt1_px': 0,
           'rgb_any_diff_px': 0, 'pixels': len(sample) * res * res}
    for t in sample:
      out = rm.Reconstruction_rotation(coeffs[t:t + 1], env.model, jit[t])                 # device reconstruction
      v_gpu, c_gpu = orc.raster_inputs(out[3], out[4], out[2], res)
      v_cpu, c_cpu, _ = pipeline.frame_raster_inputs(coeffs[t:t + 1], env.model, jit[t][0], res)
      image = np.zeros(res * res * 3, np.uint8)
      mask = np.zeros(res * res, np.uint8)
      depth = np.full(res * res, -99999.0, np.float32)
      tid_gpu = mc.render_colors_with_triangle_id(image, mask, v_gpu, tris, c_gpu, depth, tris.size // 3, res, res, 3)
      tid_cpu = np.zeros(res * res, np.int32)
      image2, mask2, depth2 = np.zeros_like(image), np.zeros_like(mask), np.full(res * res, -99999.0, np.float32)
      Oracle.render_colors(image2, mask2, v_cpu, tris, c_cpu, depth2, tris.size // 3, res, res, 3, triangle_out=tid_cpu)
      near = Oracle.near_ties(v_cpu, tris, tris.size // 3, res, res, ulps=1)
      for k, v in pipeline.classify_mismatches(v_cpu, v_gpu, tris, tid_cpu, tid_gpu, near, res).items():
        rep[k] += v
      d = np.abs(got[t].astype(np.int16) - image2.reshape(res, res, 3).astype(np.int16)).max(axis=2)
      rep['rgb_gt1_px'] += int((d > 1).sum())
      rep['rgb_any_diff_px'] += int((d > 0).sum())
    rep['sample'] = ('frames %s of the GRID utterance at %dx%d: winning triangle per pixel of the device chain against the '
                     'CPU chain (near_tie = two nearest depths within 1 ulp; edge_flip = inside test flipped on an edge of a '
                     'triangle whose corner differs in the last float32 ulp), and rendered RGB of the fused path against the '
                     'CPU frame' % (sample, res, res))
    return rep
  except Exception as e:      # the bench line must not die on the checker
    return {'error': repr(e)}


def run_sharded(env, name, steps, warmup):
  """N > 1: strong scaling of one configuration; frames land in rank 0's buffer inside the timed step."""
  torch, dist = env.torch, env.dist
  from voicepuppet_b200 import _lib, render, synthetic
  frames, res = CONFIGS[name][0], CONFIGS[name][1]
  world, rank, dev, dm, lib = env.world, env.rank, env.dev, env.dm, env.lib
  coeffs_all = synthetic.make_coeffs(frames, seed=1)
  angles_all = render.jitter_angle_sequence(frames)
  dm.set_identity(coeffs_all[0:1, :80], coeffs_all[0:1, 144:224])
  gather_mode = os.environ.get('VPB200_GATHER', 'p2p')
  peer_mode = os.environ.get('VPB200_PEER_MODE', 'auto')
  # Root-aware sharding (peer-memory gather only): rank 0 also receives everybody else's frames through one NVLink
  # port, which bounds the step from 8 ranks on; it renders a few more frames itself so that fewer cross the link.
  # R = frames/s of one GPU, measured here on rank 0; B = rank 0's ingest, measured in round 2 (profiles/r02h_*).
  root_frames, shard_model = None, None
  if gather_mode == 'p2p' and world >= 3 and os.environ.get('VPB200_ROOT_AWARE', '1') == '1':
    box = [None]
    if rank == 0:
      k = min(frames, 768)
      ex_p, par_p = render.device_inputs(coeffs_all[:k], angles_all[:k], dev)
      buf_p = torch.empty((k, res, res, 3), dtype=torch.uint8, device=dev)
      best = None
      for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        render.render_device(dm, ex_p, par_p, True, res, buf_p)
        b.record()
        torch.cuda.synchronize(dev)
        best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
      del buf_p
      fps = k / (best * 1e-3)
      gbs = float(os.environ.get('VPB200_INGEST_GBS', '720'))
      box[0] = (render.root_aware_frames(frames, world, res * res * 3, fps, gbs), fps, gbs)
    dist.broadcast_object_list(box, src=0)
    root_frames, fps, gbs = box[0]
    shard_model = {'root_frames': int(root_frames), 'render_fps_measured': round(fps), 'ingest_gbs_assumed': gbs}
  bounds = [render.shard_bounds(frames, world, r, root_frames) for r in range(world)]
  begin, end = bounds[rank]
  per = max(e - b for b, e in bounds)
  coeffs, angles = coeffs_all[begin:end], angles_all[begin:end]
  ex_dev, params_dev = render.device_inputs(coeffs, angles, dev)
  peer = render.PeerFrameBuffer(per, res, world, rank, dev, bounds=bounds) if gather_mode == 'p2p' else None
  total_slots = bounds[-1][1] if peer is not None else world * per
  local = None if peer is not None else torch.empty((per, res, res, 3), dtype=torch.uint8, device=dev)
  full = [None]

  def step():
    if peer is not None:     # finished chunks are pushed into rank 0's buffer over NVLink under the next chunk's rendering
      full[0] = peer.render_into(dm, ex_dev, params_dev, True, mode=peer_mode)
    else:                    # baseline: render locally, NCCL gather to rank 0
      full[0] = render.pipelined_gather(dm, ex_dev, params_dev, True, res, local, world, rank)

  for _ in range(max(warmup, 3)):
    env.flush_l2()
    step()
  env.barrier()
  sampler = ClockSampler(env.local_rank)
  if rank == 0:
    sampler.start()
  align = torch.zeros(1, dtype=torch.int32, device=dev)
  launches0 = lib.vp_launch_count()
  starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
  ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
  env.barrier()
  for i in range(steps):
    env.flush_l2()
    dist.all_reduce(align)           # device-side barrier: every rank's timed step starts together
    starts[i].record()
    step()
    ends[i].record()
  env.barrier()
  launches = lib.vp_launch_count() - launches0
  total_ms = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], dtype=torch.float64, device=dev)
  dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
  ms_per_step = float(total_ms.item()) / steps
  # render-only time of the shards (no gather), max over ranks: what is left of the step is the exposed gather
  r0 = torch.cuda.Event(enable_timing=True)
  r1 = torch.cuda.Event(enable_timing=True)
  scratch = torch.empty((end - begin, res, res, 3), dtype=torch.uint8, device=dev)
  render.render_device(dm, ex_dev, params_dev, True, res, scratch)
  env.barrier()
  env.flush_l2()
  r0.record()
  render.render_device(dm, ex_dev, params_dev, True, res, scratch)
  r1.record()
  torch.cuda.synchronize(dev)
  render_ms = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
  dist.all_reduce(render_ms, op=dist.ReduceOp.MAX)
  render_ms = float(render_ms.item())
  t_end = time.time() + 1.0
  while time.time() < t_end:
    render.render_device(dm, ex_dev, params_dev, True, res, scratch)
    torch.cuda.synchronize(dev)
  env.barrier()
  clocks = sampler.stop() if rank == 0 else None

  # ---- gather verification (outside the timed region): per-frame checksums of what every rank rendered locally,
  # summed over the ranks, against the same checksums of rank 0's gathered buffer after one more step
  env.flush_l2()
  step()
  env.barrier()
  weights = (torch.arange(1, res * res * 3 + 1, dtype=torch.int64, device=dev) % 8191) + 1

  def frame_sums(t):                                    # position-weighted, so that misplaced bytes do not cancel
    out = torch.empty(t.shape[0], dtype=torch.int64, device=dev)
    for a in range(0, t.shape[0], 64):
      out[a:a + 64] = (t[a:a + 64].reshape(-1, res * res * 3).to(torch.int64) * weights).sum(dim=1)
    return out
  mine = torch.zeros(frames, dtype=torch.int64, device=dev)      # contiguous shards: slot == global frame index
  mine[begin:end] = frame_sums(scratch)
  dist.all_reduce(mine)
  verified = None
  if rank == 0:
    got = frame_sums(full[0][:frames])
    verified = bool(torch.equal(got, mine)) and bool(mine.ne(0).all())
  del scratch

  # ---- end to end: host coefficient rows in, rank 0's gathered buffer drained to page-locked host memory
  host_t = None
  if rank == 0:
    host = _lib.pinned_empty((total_slots, res, res, 3), np.uint8)
    host_t = torch.from_numpy(np.asarray(host))

  def e2e_step():
    ex2, par2 = render.device_inputs(coeffs, angles, dev)       # h2d of this rank's coefficient rows
    if peer is not None:
      f = peer.render_into(dm, ex2, par2, True, mode=peer_mode)
    else:
      f = render.pipelined_gather(dm, ex2, par2, True, res, local, world, rank)
    if rank == 0:
      host_t.copy_(f[:total_slots], non_blocking=True)          # d2h of the gathered frames
    torch.cuda.synchronize(dev)
  for _ in range(2):
    e2e_step()
  env.barrier()
  n_e2e = max(3, min(steps, 10))
  t0 = time.perf_counter()
  for _ in range(n_e2e):
    dist.all_reduce(align)
    e2e_step()
  env.barrier()
  e2e_sec = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
  dist.all_reduce(e2e_sec, op=dist.ReduceOp.MAX)
  e2e_sec = float(e2e_sec.item())

  # ---- the same configuration on ONE GPU (rank 0 alone, the others idle): the strong-scaling reference point
  n1 = None
  if rank == 0 and os.environ.get('VPB200_BENCH_N1', '1') == '1':
    ex1, par1 = render.device_inputs(coeffs_all, angles_all, dev)
    buf1 = torch.empty((frames, res, res, 3), dtype=torch.uint8, device=dev)
    for _ in range(2):
      render.render_device(dm, ex1, par1, True, res, buf1)
    torch.cuda.synchronize(dev)
    k = max(3, min(steps, 5))
    s1 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
    for i in range(k):
      env.flush_l2()
      s1[i].record()
      render.render_device(dm, ex1, par1, True, res, buf1)
      e1[i].record()
    torch.cuda.synchronize(dev)
    ms1 = sum(a.elapsed_time(b) for a, b in zip(s1, e1)) / k
    n1 = {'value': frames / (ms1 * 1e-3), 'ms_per_step': ms1,
          'note': 'the same %d frames rendered by rank 0 alone in this run (no gather): the 1-GPU point of the strong scaling' % frames}
    del buf1
  env.barrier()
  if rank != 0:
    return None, peer
  _, total_bytes = algorithmic_bytes(frames, res)
  peak, _ = measured_peaks()
  eff_mode = peer.effective_mode(peer_mode, end - begin) if peer is not None else 'nccl'
  gather_label = {
      'push': 'finished chunks of frames pushed into rank 0 buffer over NVLink (CUDA IPC peer memory, copy engine) under '
              'the rendering of the next chunk (plan %s); device-side completion flags' % render.push_plan(end - begin, world),
      'store': 'resolve kernels store straight into rank 0 buffer over NVLink (CUDA IPC peer memory); device-side completion flags',
      'store-chunks': 'resolve kernels store straight into rank 0 buffer over NVLink, chunked over two streams',
      'nccl': 'NCCL gather of uint8 frames to rank 0, per chunk on a side stream'}.get(eff_mode, eff_mode)
  frame_bytes = res * res * 3
  ingest = (frames - (end - begin)) * frame_bytes            # rank 0 here: everything it did not render itself
  exposed_ms = max(0.0, ms_per_step - render_ms)
  pipeline_gbs = total_bytes / (ms_per_step * 1e-3) / 1e9
  return {
      'name': name, 'frames': frames, 'res': res, 'ms_per_step': ms_per_step, 'value': frames / (ms_per_step * 1e-3),
      'launches': int(launches), 'clocks': clocks, 'gather_verified': verified,
      'gather': {'mode': gather_label,
                 'rank0_ingest_bytes_per_step': ingest, 'render_only_ms': render_ms, 'exposed_gather_ms': exposed_ms,
                 'rank0_ingest_gbs_over_step': round(ingest / ms_per_step / 1e6, 1),
                 'completion_wait_timeouts': int(lib.vp_peer_timeouts())},
      'e2e': {'value': frames / e2e_sec, 'unit': 'frames/s',
              'h2d_bytes_per_step': frames * (64 * 4 + 192), 'd2h_bytes_per_step': total_slots * frame_bytes,
              'note': 'every rank uploads its coefficient rows, renders and pushes; rank 0 then drains the gathered '
                      'buffer to page-locked host memory (one PCIe link: the drain bounds this figure)'},
      'roofline_pipeline': {'algorithmic_bytes': total_bytes, 'achieved': round(pipeline_gbs, 1), 'peak': peak * world,
                            'unit': 'GB/s', 'frac': round(pipeline_gbs / (peak * world), 4),
                            'note': 'peak = %d x the measured single-GPU figure' % world},
      'n1_same_config': n1, 'shards': [e - b for b, e in bounds], 'shard_model': shard_model,
  }, peer


def run_ours(args):
  env = Env()
  world, rank = env.world, env.rank
  name = args.config or default_config(world)
  frames, res = CONFIGS[name][0], CONFIGS[name][1]

  if world > 1:
    r, peer = run_sharded(env, name, args.steps, args.warmup)
    # per-kernel roofline of the shard-sized launches: rank 0 alone, after the collective part is over
    if rank == 0:
      per = -(-frames // world)
      line = {
          'metric': METRIC, 'value': r['value'], 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
          'warmup': max(args.warmup, 3), 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong',
          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': {'workload': workload_name(name), 'name': name, 'frames_total': frames, 'frames_per_gpu': r['shards'],
                     'shard_model': r['shard_model'],
                     'resolution': res, 'model': MODEL_NOTE, 'coeff_seed': 1, 'l2': env.l2_note(), 'gather': r['gather']['mode'],
                     'sharding': 'contiguous frame ranges, model replicated, no data-path collective; frames gathered into rank 0 buffer '
                                 'inside the step; from 3 ranks on rank 0 renders a few more frames than the others (root-aware split: what it '
                                 'renders itself does not cross its NVLink port, which bounds the step at 8 ranks)'},
          'roofline': None, 'roofline_pipeline': r['roofline_pipeline'], 'gather': r['gather'],
          'gather_verified': r['gather_verified'], 'n1_same_config': r['n1_same_config'],
          'cpu_baseline': None, 'e2e': r['e2e'], 'gpu_launches': r['launches'], 'clocks': r['clocks'],
      }
      try:
        CONFIGS['_shard'] = (per, res, None, "one rank's shard of the frame-sharded batch: %d frames at %dx%d" % (per, res, res))
        m = measure_single_gpu(env, '_shard', max(3, min(args.steps, 5)), 3, e2e_steps=3)
        line['roofline'] = m.get('roofline')
        line['kernels'] = m['kernels']
      except Exception as e:
        line['roofline'] = {'error': repr(e)}
      print(json.dumps(line))
    env.barrier()
    if peer is not None:
      peer.close()
    env.dist.destroy_process_group()
    return

  main = measure_single_gpu(env, name, args.steps, args.warmup, want_clocks=True)
  others = {}
  if not args.no_all_configs and args.config is None:
    for other in ('single', 'grid', 'clip1500', 'sharded12000'):
      if other == name:
        continue
      try:
        m = measure_single_gpu(env, other, max(3, min(args.steps, 5)), 3, e2e_steps=3)
        others[other] = {k: m[k] for k in ('workload', 'frames', 'resolution', 'value', 'ms_per_step', 'e2e', 'e2e_clip',
                                           'kernels', 'roofline', 'roofline_pipeline', 'gpu_launches') if k in m}
      except Exception as e:
        others[other] = {'error': repr(e)}
  parity = None if args.no_cpu_baseline else tri_id_report(env)
  cpu = None
  if not args.no_cpu_baseline:
    cores = os.cpu_count() or 1
    sample = args.cpu_frames or cpu_sample_frames(frames, res, cores, 20.0)
    cpu = cpu_baseline_dict(cpu_reference_run(min(sample, frames), res, 1, 1), res)
  line = {
      'metric': METRIC, 'value': main['value'], 'unit': 'frames/s', 'n_gpus': 1, 'steps': args.steps,
      'warmup': max(args.warmup, 3), 'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': workload_name(name), 'name': name, 'frames': frames, 'resolution': res,
                 'model': MODEL_NOTE, 'coeff_seed': 1, 'l2': env.l2_note(), 'gather': 'none (one GPU)'},
      'roofline': main.get('roofline'), 'roofline_pipeline': main['roofline_pipeline'], 'kernels': main['kernels'],
      'cpu_baseline': cpu, 'e2e': main['e2e'], 'e2e_clip': main['e2e_clip'], 'gpu_launches': main['gpu_launches'],
      'clocks': main['clocks'], 'parity': parity, 'all_configs': others,
  }
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--config', default=None, choices=sorted(CONFIGS), help='default: stress4096 at N=1, sharded12000 at N>1')
  ap.add_argument('--frames', type=int, default=None, help='ad-hoc workload: frames (with --res)')
  ap.add_argument('--res', type=int, default=None)
  ap.add_argument('--cpu-frames', type=int, default=0, help='frames of the bounded CPU-baseline sample (0 = automatic)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-all-configs', action='store_true')
  args = ap.parse_args()
  if args.frames is not None or args.res is not None:     # ad-hoc size for experiments
    f, r = args.frames or 75, args.res or 256
    CONFIGS['custom'] = (f, r, None, 'custom: %d frames at %dx%d' % (f, r, r))
    args.config = 'custom'
    args.no_all_configs = True
  if args.impl == 'reference':
    run_reference_arm(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
