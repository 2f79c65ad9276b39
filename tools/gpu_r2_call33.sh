#!/bin/bash
# Round 2, call 33: resolve pass with 16 pixels per thread against 8; then everything as shipped.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call33.log) 2>&1
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "2048 1024" "1500 512" "3000 256"; do set -- $cfg; export F=$1 R=$2
  for px in 8 16; do echo "== ${F}x${R} resolve px=$px"; VPB200_RESOLVE_PX=$px b; done
done
echo "== all gpu tests (defaults)"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
