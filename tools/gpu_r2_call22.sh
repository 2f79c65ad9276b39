#!/bin/bash
# Round 2, call 22: REDG.64 vs REDG.32 lane rate (microbenchmark); scatter kernel with the next frame's vertex records prefetched.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call22.log) 2>&1
echo "== REDG microbenchmark"; timeout 120 tools/bin/diag_redg
echo "== raster tests with prefetch"; VPB200_SCATTER_PREFETCH=1 timeout 600 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -m gpu -x -q 2>&1 | tail -2
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "3000 256" "1500 512" "1024 1024"; do set -- $cfg; export F=$1 R=$2
  for pf in 0 1; do for mb in 5 4; do for fpb in 0 8; do echo "== ${F}x${R} prefetch=$pf minb=$mb fpb=$fpb"; VPB200_SCATTER_PREFETCH=$pf VPB200_SCATTER_MINB=$mb VPB200_SCATTER_FPB=$fpb b; done; done; done
done
