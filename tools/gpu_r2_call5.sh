#!/bin/bash
# Round 2, call 5: all GPU tests, K1 epilogue A/B, the default bench + reference arm end to end, ncu captures per resolution.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for e in 0 1; do VPB200_BASIS_EPI=$e timeout 120 python tools/diag_basis_ab.py; done
echo "== default bench"; t0=$(date +%s); timeout 900 python bench.py 2> gpurun_out/r02e_bench_default.err | tail -1 > gpurun_out/r02e_bench_default.json; echo "wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r02e_bench_default.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_default.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','parity')}); print({k:v['ms'] for k,v in d['kernels'].items()}); print({k:(round(v['value']),round(v['e2e']['value']),round(v['ms_per_step'],3)) for k,v in d['all_configs'].items() if 'value' in v}); print(d['all_configs'])" | cut -c1-3000
echo "== reference arm"; t0=$(date +%s); timeout 900 python bench.py --impl reference 2> gpurun_out/r02e_ref.err | tail -1 | tee gpurun_out/r02e_bench_ref.json | cut -c1-900; echo "wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r02e_ref.err
for cfg in "75 256" "128 512" "64 1024"; do
  set -- $cfg
  echo "== ncu frames=$1 res=$2"
  FRAMES=$1 RES=$2 STEPS=3 timeout 900 ncu --set full --clock-control none --import-source on \
     -k regex:"basis_tc_kernel|vertex_fan_kernel|raster_scatter_packed_kernel|resolve_packed_kernel" -s 8 -c 4 \
     -o gpurun_out/r02e_$1x$2 -f python tools/prof_step.py > gpurun_out/r02e_ncu_$1x$2.log 2>&1
  tail -2 gpurun_out/r02e_ncu_$1x$2.log
done
