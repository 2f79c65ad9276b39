"""GPU, multi-rank: contiguous frame shards on 2 GPUs, pipelined NCCL gather to rank 0
(voicepuppet_b200.render.render_sequence_sharded) == the single-GPU render of the whole sequence.
Skipped on boxes with one GPU (the gloo test in test_sharded_gloo.py covers the host logic)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, n_frames, res, queue, gather):
  sys.path.insert(0, ROOT)
  import torch
  import torch.distributed as dist
  from voicepuppet_b200 import render, synthetic
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
  try:
    model = synthetic.cached_model()
    coeffs = synthetic.make_coeffs(n_frames, seed=9)
    out = render.render_sequence_sharded(coeffs, model, res=res, angles='jitter', notify_frames=13, gather=gather)
    torch.cuda.synchronize()
    if rank == 0:
      queue.put(out.cpu().numpy())
    else:
      assert out is None
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('gather', ['p2p-push', 'p2p-store', 'nccl'])
def test_two_gpu_shards_equal_single_gpu(gather):
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  import torch.multiprocessing as mp
  from voicepuppet_b200 import render, synthetic
  n_frames, res, world = 77, 224, 2
  synthetic.cached_model()
  ctx = mp.get_context('spawn')
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, res, queue, gather)) for r in range(world)]
  for p in procs:
    p.start()
  got = queue.get(timeout=240)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  want = np.asarray(render.render_sequence(synthetic.make_coeffs(n_frames, seed=9), synthetic.cached_model(), res=res))
  assert got.shape == want.shape and np.array_equal(got, want)
