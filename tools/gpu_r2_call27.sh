#!/bin/bash
# Round 2, call 27: pairs / quads at 512x512 (forced group walk); basis store policy at long launches.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call27.log) 2>&1
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
export F=1500 R=512; echo "== ${F}x${R} defaults"; b
for g in 2 4; do for gm in 128 192 256; do for mb in 5; do echo "== ${F}x${R} forced group=$g group_min=$gm minb=$mb"; VPB200_WALK_GROUP_RES=256 VPB200_WALK_GROUP=$g VPB200_WALK_GROUP_MIN=$gm VPB200_SCATTER_MINB=$mb b; done; done; done
for sp in 0 1; do echo "== basis store policy $sp"; VPB200_BASIS_STORE=$sp timeout 300 python tools/diag_basis_blocks.py 2>&1 | tail -8; done
