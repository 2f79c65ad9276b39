#!/bin/bash
# Round 2, call 25: group walk as shipped (templated, large frames only); pairs at 512x512; 768x768 check.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call25.log) 2>&1
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
export F=3000 R=256; echo "== ${F}x${R} defaults"; b
export F=1024 R=1024; echo "== ${F}x${R} defaults"; b
export F=1500 R=512; echo "== ${F}x${R} defaults"; b
for g in 2 4; do for gm in 64 128 192; do for mb in 4 5; do echo "== ${F}x${R} forced group=$g group_min=$gm minb=$mb"; VPB200_WALK_GROUP_RES=256 VPB200_WALK_GROUP=$g VPB200_WALK_GROUP_MIN=$gm VPB200_SCATTER_MINB=$mb b; done; done; done
export F=1200 R=768; echo "== ${F}x${R} defaults"; b; echo "== ${F}x${R} group off"; VPB200_WALK_GROUP=0 b
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
