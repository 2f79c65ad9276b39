"""Host-side model of the shared-memory bank conflicts of the fan vertex kernel's position gathers
(voicepuppet_b200/csrc/reconstruct.cu, vertex_fan_kernel), computed from the topology tables alone -- no GPU.

An LDS.128 is served per quarter-warp (8 consecutive lanes); a float4 slot s occupies bank group s % 8; lanes
reading the same slot broadcast, distinct slots in one bank group serialise.  Excess wavefronts per (tile, frame)
= sum over the 9 fan entries and the 16 quarter-warps of (max distinct slots per bank group - 1).

Validated against ncu (profiles/r01z_ncu_summary.md): the model gives 128.1 excess wavefronts per tile and frame
for the synthetic full model, ncu's l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum is 2,761,124 per 75-frame
launch = 131.9 per tile and frame.  Use it to evaluate table orderings before spending GPU time:
only 12 % of the gathers hit halo slots, so re-assigning halo slots alone cannot help much; reordering the own
vertices inside a tile (slot == lane today) got 125 -> 95 with a short annealing run; decoupling slot from lane
(a per-thread slot table) would allow a proper 8-colouring."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def fan_entries(fan):
  f = fan.astype(np.int64)
  off = np.stack([f[:, 0] & 0xFFFF, f[:, 0] >> 16, f[:, 1] & 0xFFFF, f[:, 1] >> 16, f[:, 2] & 0xFFFF, f[:, 2] >> 16,
                  f[:, 3] & 0xFFFF, f[:, 3] >> 16, f[:, 4] & 0xFFFF], axis=1)
  return off // 16


def excess_wavefronts(tiles, fan):
  total, ideal, by_entry = 0, 0, np.zeros(9)
  for v_begin, nv, nlv, nlt, halo_off, ltri_off, is_fan in tiles:
    if not is_fan:
      continue
    u = fan_entries(fan[v_begin:v_begin + nv])
    for i in range(9):
      col = np.full(128, -1)
      col[:nv] = u[:, i]
      for q in range(16):
        s = col[q * 8:(q + 1) * 8]
        s = s[s >= 0]
        if len(s) == 0:
          continue
        w = np.bincount(np.unique(s) % 8, minlength=8).max()
        total += w - 1
        ideal += 1
        by_entry[i] += w - 1
  return total, ideal, by_entry


def colouring_potential(tiles, fan, step=31):
  """What decoupling slot from lane would buy: greedy 8-colouring (+ swap refinement) of the local vertices of
  every `step`-th tile so that the distinct vertices a quarter-warp reads at one fan step fall into different
  bank groups.  Returns (excess with slot == local index, excess with the colouring)."""
  before = after = 0
  rng = np.random.default_rng(1)
  for v_begin, nv, nlv, nlt, halo_off, ltri_off, is_fan in tiles[::step]:
    u = fan_entries(fan[v_begin:v_begin + nv])
    groups = []
    for i in range(9):
      for q in range(16):
        lo, hi = q * 8, min((q + 1) * 8, nv)
        if lo < nv:
          s = np.unique(u[lo:hi, i])
          if len(s) > 1:
            groups.append(s)
    cost = lambda cls: sum(np.bincount(cls[g], minlength=8).max() - 1 for g in groups)
    before += cost(np.arange(nlv) % 8)
    member = [[] for _ in range(nlv)]
    for gi, g in enumerate(groups):
      for x in g:
        member[x].append(gi)
    cap = -(-nlv // 8)
    cnt = np.zeros(8, int)
    cls = np.full(nlv, -1)
    gcount = np.zeros((len(groups), 8), int)
    for x in np.argsort([-len(mm) for mm in member]):
      best = None
      for c in range(8):
        if cnt[c] >= cap:
          continue
        add = sum(1 for gi in member[x] if gcount[gi, c] + 1 > max(1, gcount[gi].max()))
        if best is None or (add, cnt[c]) < best[0]:
          best = ((add, cnt[c]), c)
      c = best[1]
      cls[x] = c
      cnt[c] += 1
      for gi in member[x]:
        gcount[gi, c] += 1
    cur = cost(cls)
    for _ in range(2000):
      a, b = rng.integers(0, nlv, 2)
      if cls[a] == cls[b]:
        continue
      cls[a], cls[b] = cls[b], cls[a]
      c = cost(cls)
      if c <= cur:
        cur = c
      else:
        cls[a], cls[b] = cls[b], cls[a]
    after += cur
  return before, after


if __name__ == '__main__':
  import test_topology_host as tt
  from voicepuppet_b200 import synthetic
  t, _, _ = tt.build(synthetic.cached_model())
  total, ideal, by_entry = excess_wavefronts(t['tiles'], t['fan'])
  n = len(t['tiles'])
  print('tiles %d: ideal gather wavefronts per frame %d, excess %d (%.1f per tile; ncu measured 131.9)' % (n, ideal, total, total / n))
  print('excess per tile by fan entry:', np.round(by_entry / n, 1))
  if '--colour' in sys.argv:
    b, a = colouring_potential(t['tiles'], t['fan'])
    print('8-colouring of the local vertices (every 31st tile): excess %d -> %d' % (b, a))
