#!/bin/bash
# usage: tools/gpu_sweep.sh <tag> "ENV=a ENV2=b" "ENV=c" ...   (each argument one configuration; '-' = defaults)
tag=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "$@"; do
  [ "$cfg" = "-" ] && cfg="VPB200_NOP=1"
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/${tag}_sweep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
