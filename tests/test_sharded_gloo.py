"""Multi-GPU host logic on CPU: world_size-2 gloo group, contiguous frame shards, gather to rank 0
(voicepuppet_b200.render.render_sequence_sharded).  The per-rank renderer is replaced by the CPU
oracle, so this checks sharding, the global jitter indexing and the gather -- not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

from voicepuppet_b200 import render

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_every_frame_once():
  for t in (0, 1, 7, 75, 1500, 12000):
    for w in (1, 2, 4, 8):
      spans = [render.shard_bounds(t, w, r) for r in range(w)]
      assert spans[0][0] == 0 and spans[-1][1] == t
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      assert max(e - b for b, e in spans) == -(-t // w) or t == 0


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, n_frames, res, queue):
  sys.path.insert(0, ROOT)
  import torch.distributed as dist
  from oracle import pipeline
  from voicepuppet_b200 import render as r, synthetic
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    model = synthetic.make_model(420, 48)
    coeffs = synthetic.make_coeffs(n_frames, seed=9)
    out = r.render_sequence_sharded(coeffs, model, res=res, angles='jitter',
                                    render_fn=lambda c, m, res, angles: pipeline.render_sequence(c, m, res, angles))
    if rank == 0:
      queue.put(out.numpy())
    else:
      assert out is None
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('n_frames', [5, 6])
def test_two_rank_gather_equals_single_process(n_frames):
  import torch.multiprocessing as mp
  from oracle import pipeline
  from voicepuppet_b200 import synthetic
  res, world = 32, 2
  ctx = mp.get_context('spawn')
  queue = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, res, queue)) for r in range(world)]
  for p in procs:
    p.start()
  got = queue.get(timeout=180)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  model = synthetic.make_model(420, 48)
  want = pipeline.render_sequence(synthetic.make_coeffs(n_frames, seed=9), model, res, 'jitter')
  assert got.shape == want.shape and np.array_equal(got, want)
  assert want.any()


def test_push_plan_covers_every_frame(monkeypatch):
  """Chunk plans of the push gather (render.push_plan): positive chunks adding up to the shard, short first and
  last chunks for long shards, one chunk for short ones; VPB200_PUSH_PLAN overrides only when it adds up."""
  from voicepuppet_b200 import render
  monkeypatch.delenv('VPB200_PUSH_PLAN', raising=False)
  for world in (2, 4, 8):
    for n in (1, 7, 47, 48, 75, 150, 1500):
      plan = render.push_plan(n, world)
      assert sum(plan) == n and min(plan) > 0
      if n < 48:
        assert plan == [n]
      else:
        assert len(plan) >= 4 and plan[0] <= n // 4 and plan[-1] <= n // 4
  assert render.push_plan(75, 8) == [9, 29, 28, 9]
  monkeypatch.setenv('VPB200_PUSH_PLAN', '25,25,25')
  assert render.push_plan(75, 8) == [25, 25, 25]
  assert render.push_plan(76, 8) != [25, 25, 25]          # does not add up: ignored


def test_root_aware_shards_cover_every_frame_once():
  """render.root_aware_frames / shard_bounds(root_frames=...): rank 0 renders more than the equal share only when its
  NVLink ingest would bound the step (8 ranks at the measured rates), never less; shards stay contiguous and cover
  [0, T) exactly once."""
  from voicepuppet_b200 import render
  fb = 256 * 256 * 3
  for world in (2, 4):
    assert render.root_aware_frames(12000, world, fb, 595e3, 720.0) == 12000 // world
  f0 = render.root_aware_frames(12000, 8, fb, 595e3, 720.0)
  assert 1500 < f0 < 2000
  assert render.root_aware_frames(12000, 8, fb, 595e3, 1e9) == 1500          # an infinitely fast link: equal shards
  for world, root in ((8, f0), (8, None), (3, 5000), (5, 12000), (4, 0)):
    bounds = [render.shard_bounds(12000, world, r, root) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == 12000
    assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
    assert all(b >= a for a, b in bounds)
    if root is not None:
      assert bounds[0] == (0, root)
      rest = [b - a for a, b in bounds[1:]]
      assert max(rest) - min(r for r in rest if r > 0 or True) <= max(rest)   # ceil-sized, the last ones may be short
