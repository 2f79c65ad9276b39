#!/bin/bash
# Round 2, call 3: why is the fused kernel slow at 256x256?  Phase-ablation timings + one ncu --set full capture.
mkdir -p gpurun_out
b() { timeout 300 python bench.py --steps 5 --warmup 3 --frames ${F:-75} --res ${R:-256} --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; }
for dbg in 0 1 2 4 6 7; do echo "== debug=$dbg"; VPB200_FUSED_DEBUG=$dbg b; done
for minb in 4 6; do echo "== minb=$minb"; VPB200_FUSED_MINB=$minb b; done
for w in 1 3 4; do echo "== waves=$w"; VPB200_FUSED_WAVES=$w b; done
for fpb in 1 2 5 38; do echo "== fpb=$fpb"; VPB200_FUSED_FPB=$fpb b; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_tile_kernel|resolve_vcol_kernel" -s 6 -c 2 \
   -o gpurun_out/r02c_fused256 -f python tools/prof_step.py > gpurun_out/r02c_ncu_full.log 2>&1
tail -3 gpurun_out/r02c_ncu_full.log
