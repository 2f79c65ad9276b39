"""Markdown summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`) for profiles/.
usage: python tools/launch_list_md.py <launches.csv> "<command the list was taken over>" > profiles/<name>.md"""
import collections, csv, sys

path, what = sys.argv[1], sys.argv[2]
OURS = ('identity_kernel', 'frame_prep_kernel', 'basis_', 'vertex_', 'raster_scatter', 'resolve_', 'fused_tile', 'peer_', 'composite')
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
  name = r['Kernel Name'].split('(')[0].split('::')[-1][:48]
  a = agg.setdefault(name, [0, 0.0])
  a[0] += 1
  a[1] += float(r['Metric Value'])
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if any(p in k for p in OURS))
print('# ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over `%s`\n' % what)
print('Cold-cache, serialised per-launch durations: the SHARES are what to compare with the bench line, not the absolutes.')
print('"share of ours" = share among this library\'s kernels (the torch kernels are the L2 flush and checksum helpers of bench.py).\n')
print('| kernel | launches | total us | avg us | share % | share of ours % |\n|---|---|---|---|---|---|')
for k, v in agg.items():
  mine = any(p in k for p in OURS)
  print('| %s | %d | %.1f | %.2f | %.1f | %s |' % (k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, 100 * v[1] / tot,
                                                 '%.1f' % (100 * v[1] / ours) if mine else '-'))
