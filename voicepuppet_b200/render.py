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


def shard_bounds(n_frames, world_size, rank):
  """Contiguous frame range [begin, end) of ``rank``: ceil-sized shards, the last ones may be short."""
  per = -(-n_frames // world_size)
  begin = min(rank * per, n_frames)
  return begin, min(begin + per, n_frames)


def gather_groups(n_local, n_groups=None):
  """Cut a rank's frames into groups whose NCCL gather overlaps the rendering of the next group."""
  if n_groups is None:
    n_groups = 1 if n_local < 32 else min(4, n_local // 16)
  n_groups = max(1, min(n_groups, max(n_local, 1)))
  edges = [round(i * n_local / n_groups) for i in range(n_groups + 1)]
  return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def pipelined_gather(render_group, local, per, world, rank, group=None, n_groups=None):
  """Render `local` ([per,res,res,3] uint8 on this rank's GPU) group by group with
  ``render_group(a, b)`` (asynchronous on the current stream) and gather every finished group to
  rank 0 on a side stream, so that the transfer of group g runs under the rendering of group g+1.
  Returns the [world*per,res,res,3] tensor on rank 0 (frames ordered by rank), None elsewhere."""
  import torch
  import torch.distributed as dist
  full = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if rank == 0 else None
  compute = torch.cuda.current_stream(local.device)
  comm = _comm_stream(local.device)
  for a, b in gather_groups(per, n_groups):
    render_group(a, b)
    done = torch.cuda.Event()
    done.record(compute)
    comm.wait_event(done)
    with torch.cuda.stream(comm):
      dst = [full[r * per + a:r * per + b] for r in range(world)] if rank == 0 else None
      dist.gather(local[a:b], dst, dst=0, group=group)
  compute.wait_stream(comm)
  return full


_comm_streams = {}


def _comm_stream(device):
  import torch
  key = (device.type, device.index)
  if key not in _comm_streams:
    _comm_streams[key] = torch.cuda.Stream(device=device)
  return _comm_streams[key]


def render_sequence_sharded(coeffs, facemodel, res=IMG, angles='jitter', group=None, render_fn=None, n_groups=None):
  """Frames are independent, so rank r renders frames [r*ceil(T/W), (r+1)*ceil(T/W)) on its own GPU
  with no communication; the only collective is the gather of the uint8 frames to rank 0 (NCCL
  over NVLink when the group's backend is nccl, issued per group of frames on a side stream so that
  it overlaps the rendering of the next group; gloo in the CPU tests, which also substitute
  ``render_fn``).  Returns [T,res,res,3] uint8 on rank 0 (a torch tensor on the group's device)
  and None elsewhere.  The jitter sequence is a function of the global frame index, so every rank
  generates all of it and slices its shard."""
  import torch
  import torch.distributed as dist
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  coeffs = np.asarray(coeffs, dtype=np.float32)
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
  shard_angles = None if angles is None else np.asarray(angles)[begin:end]
  if use_cuda and render_fn is None:
    def render_group(a, b):
      b = min(b, end - begin)
      if b > a:
        render_sequence(coeffs[begin + a:begin + b], facemodel, res=res,
                        angles=None if shard_angles is None else shard_angles[a:b], device=device.index,
                        out=local[a:b])
    full = pipelined_gather(render_group, local, per, world, rank, group, n_groups)
    return None if rank != 0 else full[:t]
  if render_fn is None:
    raise RuntimeError('render_sequence_sharded needs CUDA ranks (nccl backend); there is no CPU path')
  if end > begin:
    frames = render_fn(coeffs[begin:end], facemodel, res=res, angles=shard_angles)
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


def render_face(center_x, center_y, ratio, bfmcoeff, img, transform_params, facemodel):
  """Same signature and result as infer_bfmvid.py:79-122.  Everything up to the rasterized 224x224
  frame runs on the GPU; the channel swap, cv2.resize and paste into the canvas are the reference's
  own cv2 calls on the host."""
  import cv2
  ratio *= transform_params[2]
  tx = -int((transform_params[3] / ratio))
  ty = -int((transform_params[4] / ratio))
  new_image = rasterize_face(bfmcoeff, facemodel, IMG)
  new_image = cv2.cvtColor(new_image, cv2.COLOR_BGR2RGB)
  new_image = cv2.resize(new_image, (int(round(new_image.shape[0] / ratio)), int(round(new_image.shape[1] / ratio))))
  back_new_image = np.zeros((img.shape[0], img.shape[1], img.shape[2]), dtype=img.dtype)
  center_face_x = new_image.shape[1] // 2
  center_face_y = new_image.shape[0] // 2
  ry = center_y - center_face_y + new_image.shape[0] - ty
  rx = center_x - center_face_x + new_image.shape[1] - tx
  back_new_image[center_y - center_face_y - ty:ry, center_x - center_face_x - tx:rx, :] = new_image
  return back_new_image
