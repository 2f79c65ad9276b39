#!/bin/bash
# Round 2, call 4: flattened big-box walk in both raster paths: parity, then separate vs fused x inline threshold.
mkdir -p gpurun_out
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_raster.py tests/test_gpu_full_sizes.py tests/test_gpu_reconstruct.py -x -q 2>&1 | tail -6
b() { timeout 600 python bench.py --steps 4 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; }
for cfg in "75 256" "1500 512" "1024 1024"; do
  set -- $cfg; export F=$1 R=$2
  for sep in 1 0; do for ib in 0 4 12 32 100000; do
    echo "== frames=$F res=$R separate=$sep inline_max=$ib"
    VPB200_BENCH_SEPARATE=$sep VPB200_INLINE_BOX=$ib b
  done; done
done
echo "== new bench, grid config"
timeout 600 python bench.py --steps 5 --warmup 3 --config grid 2>&1 | tail -1 > gpurun_out/r02d_bench_grid.json; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_grid.json')); print({k:d[k] for k in ('value','ms_per_step','parity','e2e','e2e_clip','roofline')})"
echo "== new bench, default (timing the whole run)"
/usr/bin/time -v timeout 900 python bench.py 2> gpurun_out/r02d_bench_time.log | tail -1 > gpurun_out/r02d_bench_default.json; grep -E "Elapsed|Maximum resident" gpurun_out/r02d_bench_time.log; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_default.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline')}); print({k:(round(v['value']),round(v['e2e']['value'])) for k,v in d['all_configs'].items() if 'value' in v})"
echo "== reference arm"
/usr/bin/time -v timeout 900 python bench.py --impl reference 2> gpurun_out/r02d_ref_time.log | tail -1 | tee gpurun_out/r02d_bench_ref.json | cut -c1-600; grep -E "Elapsed" gpurun_out/r02d_ref_time.log
