"""The frame loop of the reference, batched: coefficient sequence -> rasterized face frames.

Reference: voicepuppet/pixrefer/infer_bfmvid.py:76-122 (``render_face`` and its module globals
``angles`` / ``shift``) and :221-243 (coefficient tiling + the per-frame loop).

  render_sequence(coeffs[T,257], facemodel, res)   one GPU, whole sequence in one call
  render_sequence_sharded(...)                     contiguous frame shards over the ranks of a
                                                   torch.distributed group, frames gathered to rank 0
  render_face(...)                                 the reference's per-frame function, same signature
"""
import numpy as np

from .model import IMG, DeviceModel, rotation_matrices


class JitterState(object):
  """The module globals of infer_bfmvid.py:76-77 and their update rule (:85-89): all three Euler
  angles step by +-0.005 per frame, the sign flips once |angle_y| exceeds 0.03."""

  def __init__(self):
    self.angles = np.array([[0, 0, 0]], dtype=np.float32)
    self.shift = 0.005

  def step(self):
    self.angles[0][0] += self.shift
    self.angles[0][1] += self.shift
    self.angles[0][2] += self.shift
    if self.angles[0][1] > 0.03 or self.angles[0][1] < -0.03:
      self.shift = -self.shift
    return self.angles

  def sequence(self, n_frames):
    """The next ``n_frames`` values of ``angles`` as [T,3] float32 (advances the state)."""
    out = np.empty((n_frames, 3), dtype=np.float32)
    for t in range(n_frames):
      out[t] = self.step()[0]
    return out


def jitter_angle_sequence(n_frames):
  """[T,3] float32: the angles frame t of a fresh run of infer_bfmvid.py is rendered with."""
  return JitterState().sequence(n_frames)


def _identity_runs(coeffs):
  """Split [T,257] into runs of consecutive frames sharing identity and texture coefficients
  (one run for a clip: infer_bfmvid.py:223-224 tiles the identity image's coefficients)."""
  t = coeffs.shape[0]
  ident = np.concatenate([coeffs[:, :80], coeffs[:, 144:224]], axis=1)
  if t <= 1 or np.all(ident == ident[0]):
    return [(0, t)]
  change = np.any(ident[1:] != ident[:-1], axis=1)
  starts = [0] + [int(i) + 1 for i in np.nonzero(change)[0]]
  return list(zip(starts, starts[1:] + [t]))


def render_sequence(coeffs, facemodel, res=IMG, angles='jitter', want_mask=False, device=0, out=None,
                    mask_out=None):
  """Render T frames.

  coeffs  [T,257] float32 (80 id | 64 exp | 80 tex | 3 angles | 27 gamma | 3 translation).
  angles  'jitter'  : Reconstruction_rotation with the reference's jitter sequence (render_face);
          [T,3]     : Reconstruction_rotation with explicit per-frame angles;
          None      : Reconstruction with each coefficient row's own angles (dataset-prep callers,
                      datasets/make_data_from_GRID.py:516-552).
  res     output size; the reference renders 224, other sizes scale the projected x, y by res/224.
  Returns uint8 [T,res,res,3] in page-locked host memory (and the coverage mask [T,res,res]).
  """
  coeffs = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float32))
  if coeffs.ndim != 2 or coeffs.shape[1] != 257:
    raise ValueError('coeffs must be [T,257]')
  t = coeffs.shape[0]
  dm = DeviceModel.of(facemodel, device)
  if isinstance(angles, str):
    if angles != 'jitter':
      raise ValueError("angles must be 'jitter', None or an array [T,3]")
    angles = jitter_angle_sequence(t)
  rotate_first = angles is not None
  if angles is None:
    angles = coeffs[:, 224:227]
  angles = np.asarray(angles).reshape(t, 3)
  rotation = rotation_matrices(angles).reshape(t, 9)
  from . import _lib
  if out is None:
    out = _lib.pinned_empty((t, res, res, 3), np.uint8)
  if want_mask and mask_out is None:
    if hasattr(out, 'data_ptr'):
      raise ValueError('pass mask_out explicitly when rendering into device memory')
    mask_out = _lib.pinned_empty((t, res, res), np.uint8)
  for a, b in _identity_runs(coeffs):
    dm.set_identity(coeffs[a:a + 1, :80], coeffs[a:a + 1, 144:224])
    dm.render_sequence(coeffs[a:b, 80:144], rotation[a:b], coeffs[a:b, 254:257], coeffs[a:b, 227:254], res=res,
                       rotate_shape_first=rotate_first, want_mask=want_mask, out=out[a:b],
                       mask_out=None if mask_out is None else mask_out[a:b])
  return (out, mask_out) if want_mask else out


def shard_bounds(n_frames, world_size, rank, root_frames=None):
  """Contiguous frame range [begin, end) of ``rank``: ceil-sized shards, the last ones may be short.
  With ``root_frames`` rank 0 takes exactly that many frames and the others share the rest (see root_aware_frames)."""
  if root_frames is None or world_size < 2:
    per = -(-n_frames // world_size)
    begin = min(rank * per, n_frames)
    return begin, min(begin + per, n_frames)
  f0 = max(0, min(int(root_frames), n_frames))
  if rank == 0:
    return 0, f0
  per = -(-(n_frames - f0) // (world_size - 1))
  begin = min(f0 + (rank - 1) * per, n_frames)
  return begin, min(begin + per, n_frames)


def root_aware_frames(n_frames, world, frame_bytes, render_fps, ingest_gbs, startup_s=1.5e-4):
  """How many frames rank 0 should render itself when it also receives everybody else's frames.  With equal shards
  the step is bounded by rank 0's NVLink ingest from 8 ranks on ((world - 1) / world of all the bytes through one
  port: 2.06 GB of the 12000-frame configuration, 2.9 ms at the ~720 GB/s measured, against 2.5 ms of rendering);
  every frame rank 0 renders itself is a frame that does not cross the link.  Minimises
     max(f0 / R,  (n - f0) / ((world - 1) R),  startup + (n - f0) * frame_bytes / B)
  over f0 (R = frames/s one GPU renders, B = ingest bytes/s); returns the equal share when that is already optimal."""
  if world < 2 or render_fps <= 0 or ingest_gbs <= 0:
    return -(-n_frames // max(world, 1))
  equal = -(-n_frames // world)
  b = frame_bytes / (ingest_gbs * 1e9)
  best, best_t = equal, None
  for f0 in range(equal, min(n_frames, 2 * equal) + 1):
    t = max(f0 / render_fps, (n_frames - f0) / ((world - 1) * render_fps), startup_s + (n_frames - f0) * b)
    if best_t is None or t < best_t - 1e-12:
      best, best_t = f0, t
  return best


def device_inputs(coeffs, angles, device):
  """Per-frame inputs of the fused path as device tensors: expression coefficients [T,64] float32 and
  packed vp_frame_params [T,192] uint8 (rotation float64[9] | translation float32[3] | gamma float32[27])."""
  import torch
  from . import _lib
  t = coeffs.shape[0]
  params = np.zeros(t, dtype=_lib.FRAME_PARAMS_DTYPE)
  params['rotation'] = rotation_matrices(np.asarray(angles).reshape(t, 3)).reshape(t, 9)
  params['translation'] = coeffs[:, 254:257]
  params['gamma'] = coeffs[:, 227:254]
  ex_dev = torch.from_numpy(np.ascontiguousarray(coeffs[:, 80:144])).to(device)
  params_dev = torch.from_numpy(params.view(np.uint8).reshape(t, 192)).to(device)
  return ex_dev, params_dev


def render_device(dm, ex_dev, params_dev, rotate_first, res, out, mask=None, notify_frames=0, plan=None):
  """vp_render_sequence_dev(_notify / _chunks) on torch CUDA tensors, asynchronous on the current stream.
  With notify_frames > 0 (uniform chunks) or plan = [n0, n1, ...] (explicit chunk sizes adding up to T)
  returns one torch event per chunk, recorded when the chunk's frames are complete."""
  import ctypes
  import torch
  from . import _lib
  t = ex_dev.shape[0]
  stream = ctypes.c_void_p(torch.cuda.current_stream(ex_dev.device).cuda_stream)
  mask_ptr = None if mask is None else ctypes.c_void_p(mask.data_ptr())
  out_ptr = ctypes.c_void_p(out if isinstance(out, int) else out.data_ptr())   # int: a (peer) device address
  if plan is None and notify_frames > 0:
    plan = [min(notify_frames, t - a) for a in range(0, t, notify_frames)]
  if not plan:
    _lib.check(_lib.lib().vp_render_sequence_dev(dm.handle, t, ex_dev.data_ptr(), params_dev.data_ptr(),
                                                 int(bool(rotate_first)), res, out_ptr, mask_ptr, stream))
    return []
  if sum(plan) != t or min(plan) <= 0:
    raise ValueError('chunk plan %r does not add up to %d frames' % (plan, t))
  n_events = len(plan)
  events = [torch.cuda.Event() for _ in range(n_events)]
  for ev in events:
    ev.record()                      # materialises the underlying cudaEvent_t
  handles = (ctypes.c_void_p * n_events)(*[ev.cuda_event for ev in events])
  sizes = (ctypes.c_int * n_events)(*plan)
  _lib.check(_lib.lib().vp_render_sequence_dev_chunks(dm.handle, t, ex_dev.data_ptr(), params_dev.data_ptr(),
                                                      int(bool(rotate_first)), res, out_ptr, mask_ptr, stream,
                                                      sizes, n_events, handles))
  return events


def push_plan(n_local, world):
  """Chunk sizes for the push gather.  Rank 0's NVLink ingest ((world - 1) frames per rendered frame) is as
  slow as the rendering itself from 8 ranks on, so the step costs about first chunk + all pushes, or all
  renders + last push, whichever is larger: a short first and a short last chunk around large middle ones."""
  import os
  env = os.environ.get('VPB200_PUSH_PLAN')
  if env:
    plan = [int(x) for x in env.split(',') if x]
    if sum(plan) == n_local:
      return plan
  if n_local < 48:
    return [n_local]
  if n_local >= 512:
    # long shards (the 12000-frame configuration: 1500 frames per GPU at 8).  Measured at 2 GPUs
    # (profiles/r02g_multi_n2_*.json): chunks of one basis group (94 frames) cost 8 % against the unchunked
    # rendering, because every chunk is three launches with a ramp and a tail; so the chunks are 256 frames, the first
    # one cut 64 + 192 so that the first push starts early (the library contracts the first 128-frame block of the
    # basis group on its own for a planned call, the rest of the group in one launch with the second chunk), the last
    # one ending with a 64-frame chunk so that the push left exposed at the end is short.
    chunk = int(os.environ.get('VPB200_PUSH_CHUNK', '256'))
    first = min(int(os.environ.get('VPB200_PUSH_FIRST', '64')), chunk // 2)
    last = int(os.environ.get('VPB200_PUSH_LAST', '64'))
    n_full, rem = divmod(n_local, chunk)
    plan = [first, chunk - first] + [chunk] * (n_full - 1)
    if rem > last + 32:
      plan += [rem - last, last]
    elif rem > 0:
      plan += [rem]
    return plan
  edge = max(8, n_local // 8)
  mid = n_local - 2 * edge
  n_mid = max(2, -(-mid // 96))     # measured at 4 and 8 GPUs: two middle chunks beat one (profiles/r01_multigpu.txt)
  base, extra = divmod(mid, n_mid)
  return [edge] + [base + (1 if i < extra else 0) for i in range(n_mid)] + [edge]


def gather_notify_frames(n_local, world):
  """Frames per notification chunk: the gather of one chunk runs under the rendering of the next.
  One chunk (no pipelining) for short shards, where smaller launches would cost more than they hide."""
  if n_local < 256 or world < 2:   # measured: below this the extra launches and NCCL calls cost more than they hide
    return 0
  return -(-n_local // 3)


def pipelined_gather(dm, ex_dev, params_dev, rotate_first, res, local, world, rank, group=None, notify_frames=None):
  """Render this rank's frames into `local` ([per,res,res,3] uint8 on its GPU; the trailing frames of a
  short last shard stay untouched) and gather them to rank 0.  The frames are rendered in chunks and
  every finished chunk is gathered on a side stream (NCCL over NVLink), so the transfer of chunk c
  runs under the rendering of chunk c+1.  Returns the [world*per,res,res,3] tensor on rank 0 (frames
  ordered by rank), None elsewhere."""
  import torch
  import torch.distributed as dist
  per = local.shape[0]
  n = ex_dev.shape[0]
  full = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if rank == 0 else None
  if notify_frames is None:
    notify_frames = gather_notify_frames(per, world)
  compute = torch.cuda.current_stream(local.device)
  comm = _comm_stream(local.device)
  if notify_frames <= 0 or notify_frames >= per:
    if n > 0:
      render_device(dm, ex_dev, params_dev, rotate_first, res, local)
    dst = [full[r * per:(r + 1) * per] for r in range(world)] if rank == 0 else None
    dist.gather(local, dst, dst=0, group=group)
    return full
  # every rank walks the same chunk grid over `per` frames, whatever its own frame count
  events = render_device(dm, ex_dev, params_dev, rotate_first, res, local, notify_frames=notify_frames) if n > 0 else []
  for c, a in enumerate(range(0, per, notify_frames)):
    b = min(a + notify_frames, per)
    if c < len(events):
      comm.wait_event(events[c])
    else:
      comm.wait_stream(compute)
    with torch.cuda.stream(comm):
      dst = [full[r * per + a:r * per + b] for r in range(world)] if rank == 0 else None
      dist.gather(local[a:b], dst, dst=0, group=group)
  compute.wait_stream(comm)
  return full


class PeerFrameBuffer(object):
  """Rank 0's [world*per,res,res,3] uint8 frame buffer mapped into every rank of the node (CUDA IPC over
  NVLink / NVSwitch).  Each rank's resolve kernel stores its frames straight into its slice of rank 0's
  HBM, so gathering the frames is fused into the rendering: no copy kernel, no staging buffer; the only
  collective left is a one-element all-reduce that orders rank 0's consumer after every rank's stores."""

  def __init__(self, per, res, world, rank, device, group=None, bounds=None):
    """per: frames per rank (rank r's slice starts at frame r * per; the last shard may be padded) -- or, with
    ``bounds`` = [(begin, end)] * world (contiguous, covering [0, T)), uneven shards in a dense [T,res,res,3] buffer."""
    import ctypes
    import torch
    import torch.distributed as dist
    from . import _lib
    if bounds is None:
      bounds = [(r * per, (r + 1) * per) for r in range(world)]
    if len(bounds) != world or bounds[0][0] != 0 or any(bounds[r][1] != bounds[r + 1][0] for r in range(world - 1)):
      raise ValueError('bounds must be contiguous frame ranges, one per rank, starting at 0')
    self.bounds = [(int(a), int(b)) for a, b in bounds]
    total = self.bounds[-1][1]
    self.per = self.bounds[rank][1] - self.bounds[rank][0]     # frames this rank contributes
    self.res, self.world, self.rank, self.group = res, world, rank, group
    self.frame_bytes = res * res * 3
    self.full = None
    self._base = None
    payload = [None]
    n_bytes = total * self.frame_bytes
    self.device = device
    self.step = 0
    self.local = None
    if rank == 0:
      # frames, then one 32-bit completion flag per rank (128-byte aligned)
      self._storage = torch.zeros(((n_bytes + 127) // 128) * 128 + 128, dtype=torch.uint8, device=device)
      self.full = self._storage[:n_bytes].view(total, res, res, 3)
      handle = (ctypes.c_ubyte * 64)()
      offset = ctypes.c_ulonglong()
      _lib.check(_lib.lib().vp_ipc_export(ctypes.c_void_p(self._storage.data_ptr()), handle, ctypes.byref(offset)))
      payload = [(bytes(handle), int(offset.value))]
      torch.cuda.synchronize(device)           # the zeroed flags are in memory before anybody signals
    dist.broadcast_object_list(payload, src=0, group=group)
    handle_bytes, offset = payload[0]
    if rank == 0:
      start = self._storage.data_ptr()
    else:
      base = ctypes.c_void_p()
      buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle_bytes)
      _lib.check(_lib.lib().vp_ipc_open(buf, device.index, ctypes.byref(base)))
      self._base = base
      start = base.value + offset
    self.slice_ptr = start + self.bounds[rank][0] * self.frame_bytes
    self.flags_ptr = start + ((n_bytes + 127) // 128) * 128
    if world > 32:
      raise ValueError('PeerFrameBuffer supports up to 32 ranks')
    dist.barrier(group)          # every rank has mapped the buffer before anybody renders, signals or waits

  def check(self):
    """After synchronising: raise if a completion wait on this device gave up (a rank never signalled)."""
    from . import _lib
    n = _lib.lib().vp_peer_timeouts()
    if n != 0:
      raise _lib.VpError('peer gather: %d completion wait(s) timed out (VPB200_PEER_TIMEOUT_S); frames are incomplete' % n)

  def effective_mode(self, mode, n_frames):
    """What mode='auto' resolves to for a shard of n_frames frames."""
    if mode != 'auto':
      return mode
    return 'store' if (self.world <= 2 and n_frames < 256) else 'push'

  def render_into(self, dm, ex_dev, params_dev, rotate_first, mode='auto', notify_frames=None):
    """Render this rank's frames and land them in its slice of rank 0's buffer, then publish a
    completion flag; on rank 0 the current stream then waits (on the device) until every rank's flag
    has arrived, so work enqueued after this call sees all the frames.
      mode='push'  : render into a local staging buffer in chunks; every finished chunk is pushed to
                     rank 0 by the copy engine on a side stream while the SMs render the next chunk
      mode='store' : the resolve kernel stores straight into the peer-mapped slice (no staging, no
                     copy; the kernel then runs at NVLink ingest speed)
      mode='auto'  : 'store' for two ranks with short shards, 'push' otherwise (measured on B200: 231 vs 261 us per
                     75-frame step at 2 GPUs, 303 vs 272 us at 4, profiles/r01_multigpu.txt; 6000-frame shards at
                     2 GPUs: push 10.98 ms, store 13.3 ms per step, profiles/r02g_multi_n2_*.json)
    Returns the full buffer on rank 0, None elsewhere."""
    import ctypes
    import torch
    from . import _lib
    lib = _lib.lib()
    mode = self.effective_mode(mode, ex_dev.shape[0])
    self.step += 1
    n = ex_dev.shape[0]
    compute = torch.cuda.current_stream(self.device)
    cstream = ctypes.c_void_p(compute.cuda_stream)
    flag = ctypes.c_void_p(self.flags_ptr + 4 * self.rank)
    if self.rank == 0 or mode in ('store', 'store-chunks'):
      if n > 0:
        # 'store-chunks': the same direct stores, cut into the push plan's chunks on the two streams of the
        # chunk pipeline, so that one chunk's resolve kernel waits on NVLink while the next chunk computes
        plan = push_plan(n, self.world) if (mode == 'store-chunks' and self.rank != 0) else None
        render_device(dm, ex_dev, params_dev, rotate_first, self.res, int(self.slice_ptr), plan=plan)
      _lib.check(lib.vp_peer_signal(flag, self.step, cstream))
    else:
      if self.local is None:
        self.local = torch.empty((self.per, self.res, self.res, 3), dtype=torch.uint8, device=self.device)
      plan = push_plan(n, self.world) if notify_frames is None else [min(notify_frames, n - a) for a in range(0, n, notify_frames)]
      copy = _comm_stream(self.device)
      sstream = ctypes.c_void_p(copy.cuda_stream)
      if n > 0:
        events = render_device(dm, ex_dev, params_dev, rotate_first, self.res, self.local, plan=plan)
        a = 0
        for c, ev in enumerate(events):
          b = a + plan[c]
          copy.wait_event(ev)
          _lib.check(lib.vp_copy_async(ctypes.c_void_p(self.slice_ptr + a * self.frame_bytes),
                                       ctypes.c_void_p(self.local.data_ptr() + a * self.frame_bytes),
                                       (b - a) * self.frame_bytes, sstream))
          a = b
      else:
        copy.wait_stream(compute)
      _lib.check(lib.vp_peer_signal(flag, self.step, sstream))
      compute.wait_stream(copy)      # the staging buffer is free again for the next call
    if self.rank == 0:
      _lib.check(lib.vp_peer_wait(ctypes.c_void_p(self.flags_ptr), self.world, self.step, cstream))
    return self.full

  def close(self):
    from . import _lib
    if self._base is not None:
      _lib.lib().vp_ipc_close(self._base)
      self._base = None


_comm_streams = {}


def _comm_stream(device):
  import torch
  key = (device.type, device.index)
  if key not in _comm_streams:
    _comm_streams[key] = torch.cuda.Stream(device=device)
  return _comm_streams[key]


def render_sequence_sharded(coeffs, facemodel, res=IMG, angles='jitter', group=None, render_fn=None,
                            notify_frames=None, gather='p2p'):
  """Frames are independent, so rank r renders frames [r*ceil(T/W), (r+1)*ceil(T/W)) on its own GPU
  with no communication; the only collective is the gather of the uint8 frames to rank 0 (NCCL
  over NVLink when the group's backend is nccl, issued per chunk of frames on a side stream so that
  it overlaps the rendering of the next chunk; gloo in the CPU tests, which also substitute
  ``render_fn``).  Returns [T,res,res,3] uint8 on rank 0 (a torch tensor on the group's device)
  and None elsewhere.  The jitter sequence is a function of the global frame index, so every rank
  generates all of it and slices its shard.  One identity per call (a clip).
  gather='p2p' (default): the frames are stored straight into rank 0's buffer over NVLink by the resolve
  kernel (PeerFrameBuffer); gather='nccl': rendered locally, then gathered with NCCL."""
  import torch
  import torch.distributed as dist
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  coeffs = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float32))
  t = coeffs.shape[0]
  if isinstance(angles, str):
    if angles != 'jitter':
      raise ValueError("angles must be 'jitter', None or an array [T,3]")
    angles = jitter_angle_sequence(t)
  begin, end = shard_bounds(t, world, rank)
  per = -(-t // world)
  use_cuda = dist.get_backend(group) == 'nccl'
  device = torch.device('cuda', torch.cuda.current_device()) if use_cuda else torch.device('cpu')
  local = torch.zeros((per, res, res, 3), dtype=torch.uint8, device=device)
  rotate_first = angles is not None
  shard_angles = coeffs[begin:end, 224:227] if angles is None else np.asarray(angles)[begin:end]
  if use_cuda and render_fn is None:
    if len(_identity_runs(coeffs)) != 1:
      raise ValueError('render_sequence_sharded renders one clip: identity and texture coefficients must not vary')
    dm = DeviceModel.of(facemodel, device.index)
    dm.set_identity(coeffs[0:1, :80], coeffs[0:1, 144:224])
    ex_dev, params_dev = device_inputs(coeffs[begin:end], shard_angles, device)
    if gather in ('p2p', 'p2p-store', 'p2p-push'):
      buf = PeerFrameBuffer(per, res, world, rank, device, group)
      full = buf.render_into(dm, ex_dev, params_dev, rotate_first, mode='store' if gather == 'p2p-store' else ('push' if gather == 'p2p-push' else 'auto'),
                             notify_frames=notify_frames)
      torch.cuda.current_stream(device).synchronize()
      buf.check()
      dist.barrier(group)          # nobody unmaps before every rank's stores are complete
      buf.close()
    else:
      full = pipelined_gather(dm, ex_dev, params_dev, rotate_first, res, local, world, rank, group, notify_frames)
    return None if rank != 0 else full[:t]
  if render_fn is None:
    raise RuntimeError('render_sequence_sharded needs CUDA ranks (nccl backend); there is no CPU path')
  if end > begin:
    frames = render_fn(coeffs[begin:end], facemodel, res=res, angles=None if angles is None else shard_angles)
    local[:end - begin].copy_(torch.from_numpy(np.ascontiguousarray(frames)))
  gathered = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
  dist.gather(local, gathered, dst=0, group=group)
  if rank != 0:
    return None
  return torch.cat(gathered, dim=0)[:t]


# ---------------------------------------------------------------------------------------------
# the reference's per-frame function, kept for scripts that call it frame by frame
# ---------------------------------------------------------------------------------------------
_state = JitterState()


def reset_jitter():
  global _state
  _state = JitterState()


def rasterize_face(bfmcoeff, facemodel, res=IMG):
  """infer_bfmvid.py:85-109: advance the jitter, reconstruct, rasterize -> uint8 [res,res,3] (RGB as
  written by render_colors_core, before the reference's channel swap)."""
  ang = _state.step().copy()
  return np.asarray(render_sequence(np.asarray(bfmcoeff).reshape(1, 257), facemodel, res=res, angles=ang)[0])


def composite_placement(center_x, center_y, ratio, transform_params, res=IMG):
  """infer_bfmvid.py:80-82,112-121: (S, x0, y0) -- side of the resized face and its top-left corner in the canvas."""
  import ctypes
  from . import _lib
  tp = np.ascontiguousarray(np.asarray(transform_params, dtype=np.float64).reshape(-1)[:5])
  if tp.size != 5:
    raise ValueError('transform_params must hold 5 values (w0, h0, s, tx, ty)')
  size, x0, y0 = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
  _lib.check(_lib.lib().vp_composite_placement(int(res), int(center_x), int(center_y), float(ratio), _lib.ptr(tp),
                                               ctypes.byref(size), ctypes.byref(x0), ctypes.byref(y0)))
  return size.value, x0.value, y0.value


def composite_device(frames, center_x, center_y, ratio, transform_params, canvas_hw=(512, 512), inputs=None,
                     channel_offset=3, want_canvas=True):
  """The post-raster part of render_face and of the frame loop (infer_bfmvid.py:111-121, 234-236) for a batch
  of rasterized frames, on the GPU: cv2.resize-exact bilinear resize, paste into a zero canvas, and the
  float32 / 255 image PixReferNet reads.

  frames   torch uint8 [T,res,res,3] on a CUDA device (as written by the rasterizer)
  inputs   optional torch float32 [T,H,W,C] on the same device; channels channel_offset..+2 are overwritten
           (the frame loop's ``inputs[0, ..., 3:6] = face3d``), the others are left alone
  Returns (canvas uint8 [T,H,W,3] with render_face's channel order, or None; inputs or None).
  Asynchronous on the current torch stream."""
  import ctypes
  import torch
  from . import _lib
  if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
          and frames.shape[1] == frames.shape[2] and frames.is_contiguous()):
    raise ValueError('frames must be a contiguous uint8 CUDA tensor [T,res,res,3]')
  t, res = int(frames.shape[0]), int(frames.shape[1])
  h, w = int(canvas_hw[0]), int(canvas_hw[1])
  size, x0, y0 = composite_placement(center_x, center_y, ratio, transform_params, res)
  canvas = torch.empty((t, h, w, 3), dtype=torch.uint8, device=frames.device) if want_canvas else None
  in_c = 0
  if inputs is not None:
    if not (inputs.is_cuda and inputs.dtype == torch.float32 and inputs.is_contiguous() and inputs.dim() == 4
            and tuple(inputs.shape[:3]) == (t, h, w) and inputs.device == frames.device):
      raise ValueError('inputs must be a contiguous float32 CUDA tensor [T,%d,%d,C] on the frames\' device' % (h, w))
    in_c = int(inputs.shape[3])
  if canvas is None and inputs is None:
    raise ValueError('nothing to produce: pass inputs or want_canvas=True')
  stream = ctypes.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)
  _lib.check(_lib.lib().vp_composite_dev(
      ctypes.c_void_p(frames.data_ptr()), t, res, size, x0, y0, h, w,
      None if canvas is None else ctypes.c_void_p(canvas.data_ptr()), 1,
      None if inputs is None else ctypes.c_void_p(inputs.data_ptr()), in_c, int(channel_offset),
      frames.device.index, stream))
  return canvas, inputs


def render_face_sequence(center_x, center_y, ratio, coeffs, img_shape, transform_params, facemodel, inputs=None,
                         channel_offset=3, angles='jitter', device=0, want_canvas=True):
  """The frame loop of infer_bfmvid.py:231-236 for a whole coefficient sequence, entirely on the GPU:
  reconstruct + rasterize T frames at 224x224 (render_face, :79-109), resize / paste them into the
  identity image's canvas (:111-121) and write face3d = canvas / 255 into ``inputs[..., 3:6]`` -- the tensor
  PixReferNet is fed with -- without a host round trip.  Returns (canvas [T,H,W,3] uint8 CUDA tensor with
  render_face's channel order or None, inputs)."""
  import torch
  coeffs = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float32))
  t = coeffs.shape[0]
  dev = torch.device('cuda', device)
  frames = torch.empty((t, IMG, IMG, 3), dtype=torch.uint8, device=dev)
  with torch.cuda.device(dev):
    render_sequence(coeffs, facemodel, res=IMG, angles=angles, device=device, out=frames)
    return composite_device(frames, center_x, center_y, ratio, transform_params, (img_shape[0], img_shape[1]),
                            inputs, channel_offset, want_canvas)


def _render_face_one(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel, angles):
  """One frame through the GPU path: reconstruction + rasterization at 224x224, channel swap, cv2.resize-exact
  resize and paste; returns the canvas as a numpy array of ``img``'s shape and dtype."""
  import torch
  dev = torch.device('cuda', 0)
  frames = torch.empty((1, IMG, IMG, 3), dtype=torch.uint8, device=dev)
  with torch.cuda.device(dev):
    render_sequence(np.asarray(bfmcoeff).reshape(1, 257), facemodel, res=IMG, angles=angles, device=0, out=frames)
    try:
      canvas, _ = composite_device(frames, center_x, center_y, ratio, transform_params, (img.shape[0], img.shape[1]))
    except Exception as e:   # numpy raises ValueError when the face does not fit the canvas (:121)
      raise ValueError(str(e))
  return canvas[0].cpu().numpy().astype(img.dtype, copy=False)


def render_face(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel):
  """Same signature and result as voicepuppet/pixrefer/infer_bfmvid.py:79-122 (one frame, jitter globals
  advanced, Reconstruction_rotation with the jitter angles): reconstruction, rasterization, channel swap,
  cv2.resize-exact resize and paste all run on the GPU; the returned canvas is a numpy uint8 array of ``img``'s
  shape, like the reference's."""
  ang = _state.step().copy()
  return _render_face_one(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel, ang)


def render_face_pixflow(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel):
  """The copy of render_face in voicepuppet/pixflow/infer_bfm_pixflow.py:72-115: the jitter update is commented
  out there (:78-82), so Reconstruction_rotation always gets the module's initial angles [[0, 0, 0]]."""
  return _render_face_one(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel,
                          np.zeros((1, 3), dtype=np.float32))


def render_face_dataset(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel):
  """The copy of render_face in datasets/make_data_from_GRID.py:516-552 (dataset preparation): Reconstruction
  with the coefficient row's own angles, single rotation."""
  return _render_face_one(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel, None)
