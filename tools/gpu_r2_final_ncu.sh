#!/bin/bash
# Final ncu evidence of round 2: launch lists of bench.py per configuration + one --set full capture of the four hot kernels
# in the launch shapes each configuration uses.  Reports stay in gpurun_out/, summaries are generated into profiles/ afterwards.
mkdir -p gpurun_out
PAT='basis_tc_kernel|basis_simt_tma_kernel|vertex_fan_kernel|raster_scatter_packed_kernel|resolve_packed_kernel'
echo "== launch list: bench.py --config grid"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches_grid.csv \
   python bench.py --config grid --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_launches_grid.log 2>&1; tail -1 gpurun_out/r02z_launches_grid.csv | cut -c1-200
echo "== launch list: bench.py (stress4096), first 1500 launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02z_launches_stress4096.csv \
   python bench.py --config stress4096 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_launches_stress.log 2>&1; tail -1 gpurun_out/r02z_launches_stress4096.csv | cut -c1-200
cap() {  # frames res skip count tag
  FRAMES=$1 RES=$2 STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$PAT" -s $3 -c $4 \
     -o gpurun_out/r02z_$5 -f python tools/prof_step.py > gpurun_out/r02z_ncu_$5.log 2>&1; tail -1 gpurun_out/r02z_ncu_$5.log
}
echo "== full captures"
cap 1 256 4 4 1x256
cap 75 256 7 7 75x256
cap 1500 512 50 8 1500x512
cap 1500 256 30 8 1500x256
cap 4096 1024 365 8 4096x1024
ls -la gpurun_out/r02z_*.ncu-rep
