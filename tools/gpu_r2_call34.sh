#!/bin/bash
# Round 2, call 34: frames per scatter CTA beyond 8 (small frames).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call34.log) 2>&1
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "3000 256" "1500 512"; do set -- $cfg; export F=$1 R=$2
  for fpb in 8 12 16 24; do echo "== ${F}x${R} scatter fpb=$fpb"; VPB200_SCATTER_FPB=$fpb b; done
done
