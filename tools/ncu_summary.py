"""Markdown summary of an `ncu --set full` report (+ optional launch list) for profiles/.
usage: python tools/ncu_summary.py <report.ncu-rep> [launches.csv] > profiles/<name>.md"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=float('nan')):
  try: return float(r[idx[k]])
  except Exception: return d
def scale(r, k):   # bytes metrics come with a unit column
  u = units[idx[k]]
  return g(r, k) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
print('# ncu --set full --clock-control none: %s\n' % rep.split('/')[-1])
print('| kernel | grid x block | time us | DRAM read MB | DRAM write MB | DRAM % | SM % | issue active % | warps active % | regs | L2 hit % | smem bank conflicts | warp-instr | tensor pipe % |')
print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
for r in data:
  name = r[idx['Kernel Name']].split('(')[0].split('::')[-1][:40]
  print('| %s | %s x %s | %.1f | %.2f | %.2f | %.1f | %.1f | %.1f | %.1f | %d | %.1f | %d | %d | %.1f |' % (
      name, r[idx['Grid Size']], r[idx['Block Size']], g(r, 'gpu__time_duration.sum'),
      scale(r, 'dram__bytes_read.sum') / 1e6, scale(r, 'dram__bytes_write.sum') / 1e6,
      g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'), g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
      g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
      g(r, 'launch__registers_per_thread'), g(r, 'lts__t_sector_hit_rate.pct'),
      g(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 0), g(r, 'smsp__inst_executed.sum'),
      g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0)))
print('\nWarp stall reasons (cycles per issued instruction; `selected` = 1 is the issue itself):\n')
stalls = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
short = lambda h: h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')
keep = [h for h in stalls if max(g(r, h, 0) for r in data) >= 0.5 and short(h) != 'selected']
print('| kernel | ' + ' | '.join(short(h) for h in keep) + ' |')
print('|---|' + '---|' * len(keep))
for r in data:
  print('| %s | ' % r[idx['Kernel Name']].split('(')[0].split('::')[-1][:40] + ' | '.join('%.2f' % g(r, h, 0) for h in keep) + ' |')
KERNEL_SLOTS = (('basis_tc_kernel', 'basis'), ('basis_simt_kernel', 'basis'), ('vertex_fan_kernel', 'vertex'),
                ('vertex_tile_kernel', 'vertex'), ('raster_scatter', 'scatter'), ('resolve_packed_kernel', 'resolve'),
                ('fused_tile_kernel', 'fused'), ('resolve_vcol_kernel', 'resolve'))
if '--traffic-json' in sys.argv:   # per-launch DRAM traffic of the hot kernels, read back by bench.py
  import json, os
  out_path = sys.argv[sys.argv.index('--traffic-json') + 1]
  key = sys.argv[sys.argv.index('--key') + 1]          # "<frames>x<res>" of the bench configuration the capture belongs to
  traffic = {'source': rep.split('/')[-1], 'note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches of the kernel), ncu --set full --clock-control none'}
  sums, counts = {}, {}
  for r in data:
    for pat, slot in KERNEL_SLOTS:
      if pat in r[idx['Kernel Name']]:
        sums[slot] = sums.get(slot, 0) + scale(r, 'dram__bytes_read.sum') + scale(r, 'dram__bytes_write.sum')
        counts[slot] = counts.get(slot, 0) + 1
  for slot in sums:
    traffic[slot] = int(sums[slot] / counts[slot])
  allt = json.load(open(out_path)) if os.path.exists(out_path) else {}
  if not all(isinstance(v, dict) for v in allt.values()):
    allt = {}                                            # round-1 layout (one flat dict): start over
  allt[key] = traffic
  json.dump(allt, open(out_path, 'w'), indent=1)
skip = set()
for flag in ('--traffic-json', '--key'):
  if flag in sys.argv:
    skip.add(sys.argv[sys.argv.index(flag) + 1])
args = [a for a in sys.argv[2:] if not a.startswith('--') and a not in skip]
if args:
  sys.argv = sys.argv[:2] + args
  lr = list(csv.DictReader(l for l in open(sys.argv[2]) if l.startswith('"')))
  agg = collections.OrderedDict()
  for r in lr:
    a = agg.setdefault(r['Kernel Name'].split('(')[0].split('::')[-1][:48], [0, 0.0, r['Grid Size']])
    a[0] += 1; a[1] += float(r['Metric Value'])
  tot = sum(v[1] for v in agg.values())
  print('\n## Launch list (%s: gpu__time_duration.sum, cold cache, serialised; shares, not absolutes)\n' % sys.argv[2].split('/')[-1])
  print('| kernel | launches | total us | avg us | share % |\n|---|---|---|---|---|')
  for k, v in agg.items():
    print('| %s | %d | %.1f | %.2f | %.1f |' % (k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, 100 * v[1] / tot))
