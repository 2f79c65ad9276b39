#!/bin/bash
# Round 2, call 21: conservative FMA filter in front of the exact inside test (scatter kernel) -- parity, then A/B.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call21.log) 2>&1
echo "== raster / sequence / full-size tests"; timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -m gpu -x -q 2>&1 | tail -3
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512" "3000 256"; do set -- $cfg; export F=$1 R=$2
  for fm in 1000000 8 4 16; do for mb in 5 4; do echo "== ${F}x${R} filter_min=$fm minb=$mb"; VPB200_FILTER_MIN=$fm VPB200_SCATTER_MINB=$mb b; done; done
done
