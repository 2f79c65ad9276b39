#!/bin/bash
# Round 2, call 26: resolve pass skips the keys outside the frame's bounding box (written by the scatter kernel); pairs at 512x512.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call26.log) 2>&1
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac']) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512" "3000 256" "75 256"; do set -- $cfg; export F=$1 R=$2
  for bb in 0 1; do echo "== ${F}x${R} resolve_bbox=$bb"; VPB200_RESOLVE_BBOX=$bb b; done
done
export F=1500 R=512
for g in 2 4; do for gm in 96 160; do for mb in 4 5; do echo "== ${F}x${R} forced group=$g group_min=$gm minb=$mb"; VPB200_WALK_GROUP_RES=256 VPB200_WALK_GROUP=$g VPB200_WALK_GROUP_MIN=$gm VPB200_SCATTER_MINB=$mb b; done; done; done
