#!/bin/bash
# Round 2, call 30: FMA filter inside the group walk (1024x1024 is issue-bound there) -- parity, then A/B against the recorded state.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call30.log) 2>&1
echo "== raster / sequence / full-size tests"; timeout 900 python -m pytest tests/test_gpu_raster.py tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -m gpu -x -q 2>&1 | tail -2
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
export F=1024 R=1024; echo "== ${F}x${R}"; b; echo "== ${F}x${R} group_min 160"; VPB200_WALK_GROUP_MIN=160 b; echo "== ${F}x${R} group 8"; VPB200_WALK_GROUP=8 b
export F=1200 R=768; echo "== ${F}x${R}"; b; echo "== ${F}x${R} group_min 160"; VPB200_WALK_GROUP_MIN=160 b
export F=1500 R=512; echo "== ${F}x${R} forced group 4, group_min 160"; VPB200_WALK_GROUP_RES=256 VPB200_WALK_GROUP_MIN=160 b; echo "== ${F}x${R} forced group 4, group_min 256"; VPB200_WALK_GROUP_RES=256 VPB200_WALK_GROUP_MIN=256 b
