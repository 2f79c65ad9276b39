#!/bin/bash
# Round 2, call 28: scatter kernel CTA size (256 / 128 / 64 threads) and frames per CTA, all resolutions.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call28.log) 2>&1
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512" "3000 256" "75 256"; do set -- $cfg; export F=$1 R=$2
  for bs in 256 128 64; do for fpb in 0 2 8; do echo "== ${F}x${R} block=$bs fpb=$fpb"; VPB200_SCATTER_BLOCK=$bs VPB200_SCATTER_FPB=$fpb b; done; done
done
echo "== raster tests, block 128"; VPB200_SCATTER_BLOCK=128 timeout 600 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -m gpu -x -q 2>&1 | tail -2
