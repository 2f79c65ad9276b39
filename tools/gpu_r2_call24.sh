#!/bin/bash
# Round 2, call 24: group walk only for warps whose boxes hold at least group_min pixels -- threshold sweep.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call24.log) 2>&1
echo "== raster / sequence / full-size tests (defaults)"; timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -m gpu -x -q 2>&1 | tail -2
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512"; do set -- $cfg; export F=$1 R=$2
  for gm in 96 160 240 320 480 640; do echo "== ${F}x${R} group=4 group_min=$gm"; VPB200_WALK_GROUP_MIN=$gm b; done
  echo "== ${F}x${R} group=8 group_min=320"; VPB200_WALK_GROUP=8 VPB200_WALK_GROUP_MIN=320 b
done
export F=3000 R=256; echo "== ${F}x${R} defaults"; b
export F=75 R=256; echo "== ${F}x${R} defaults"; b
