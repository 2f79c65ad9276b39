#!/bin/bash
# Final single-GPU numbers of round 2, as the driver runs them.
mkdir -p gpurun_out
t0=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02_final_ref.err | tail -1 > gpurun_out/r02_final_bench_ref.json; echo "reference arm wall $(( $(date +%s) - t0 )) s"; cut -c1-300 gpurun_out/r02_final_bench_ref.json
t0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02_final.err | tail -1 > gpurun_out/r02_final_bench.json; echo "bench wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/r02_final.err
python -c "
import json; d=json.load(open('gpurun_out/r02_final_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_clip','roofline','cpu_baseline','parity','gpu_launches','clocks')})
print({k:(v['ms'],v['frac'],v['launches']) for k,v in d['kernels'].items()})
for k,v in d['all_configs'].items(): print(k, round(v['value']), round(v['e2e']['value']), round(v['e2e_clip']['value']), round(v['ms_per_step'],4), {a:(b['ms'],b['frac']) for a,b in v['kernels'].items()}, v['roofline']['kernel'], v['roofline']['frac'], v['roofline']['traffic'])
" | cut -c1-2500
