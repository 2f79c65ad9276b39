#!/bin/bash
# Re-capture after the 8-pixel resolve: the stress and clip launch shapes (ncu --set full), and the stress launch list.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02r_ncu.log) 2>&1
PAT='basis_tc_kernel|basis_simt_tma_kernel|vertex_fan_kernel|raster_scatter_packed_kernel|resolve_packed_kernel'
cap() {  # frames res steps skip count tag
  FRAMES=$1 RES=$2 STEPS=$3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s $4 -c $5 \
     -o gpurun_out/r02r_$6 -f python tools/prof_step.py > gpurun_out/r02r_ncu_$6.log 2>&1; tail -1 gpurun_out/r02r_ncu_$6.log
}
cap 4096 1024 2 313 7 4096x1024
cap 1500 512 3 88 8 1500x512
echo "== launch list: bench.py (stress4096), first 1500 launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02r_launches_stress4096.csv \
   python bench.py --config stress4096 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_launches_stress.log 2>&1; tail -1 gpurun_out/r02r_launches_stress4096.csv | cut -c1-160
