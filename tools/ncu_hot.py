"""Summarise an ncu source page (SASS + sampling) of one kernel: top instructions by stall samples and by issue count.
usage: python tools/ncu_hot.py <report.ncu-rep> <kernel regex> [top]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '-k', 'regex:' + pat], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[1:] if len(r) == len(hdr)]
def f(r, k):
  try: return float(r[idx[k]])
  except: return 0.0
tot_s = sum(f(r, '# Samples') for r in body); tot_i = sum(f(r, 'Instructions Executed') for r in body)
print('instructions: %d SASS lines, %.0f warp-instr executed, %.0f samples' % (len(body), tot_i, tot_s))
stall_cols = [h for h in hdr if h.startswith('stall_')]
print('stall totals:', {h[6:]: int(sum(f(r, h) for r in body)) for h in stall_cols if sum(f(r, h) for r in body) > 0.01 * tot_s})
print('--- by samples')
for n, r in sorted(enumerate(body), key=lambda t: -f(t[1], '# Samples'))[:top]:
  st = {h[6:]: int(f(r, h)) for h in stall_cols if f(r, h) >= 0.2 * max(1.0, f(r, '# Samples'))}
  print('%4d %5.1f%% smp  %5.2f%% inst  %-70s %s' % (n, 100 * f(r, '# Samples') / tot_s, 100 * f(r, 'Instructions Executed') / tot_i, r[idx['Source']][:70], st))
if '--sass' in sys.argv:
  print('--- full listing')
  for n, r in enumerate(body):
    print('%4d %5.1f%% %5.2f%% w=%s/%s %s' % (n, 100 * f(r, '# Samples') / tot_s, 100 * f(r, 'Instructions Executed') / tot_i, r[idx['L1 Wavefronts Shared']], r[idx['L1 Wavefronts Shared Ideal']], r[idx['Source']][:90]))
