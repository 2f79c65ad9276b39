#!/bin/bash
# Round 2, call 23: group walk (neighbouring lanes take neighbouring pixels of one box: coalesced REDG) -- parity, then A/B.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call23.log) 2>&1
for g in 4 8; do echo "== raster / sequence / full-size tests, group $g"; VPB200_WALK_GROUP=$g timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -m gpu -x -q 2>&1 | tail -2; done
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512" "3000 256" "75 256"; do set -- $cfg; export F=$1 R=$2
  for g in 0 4 8 16; do for mb in 5 4; do echo "== ${F}x${R} group=$g minb=$mb"; VPB200_WALK_GROUP=$g VPB200_SCATTER_MINB=$mb b; done; done
done
