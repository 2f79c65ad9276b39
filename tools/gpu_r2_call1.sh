#!/bin/bash
# Round 2, call 1: GPU tests, the opt-in flavours prepared in round 1 (A/B), and first numbers on the other BASELINE configs.
mkdir -p gpurun_out
free -g | head -2; nproc; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/gpu_round2_experiments.sh 2>&1
for cfg in "1 256" "1500 512" "4096 1024"; do
  set -- $cfg
  echo "== bench frames=$1 res=$2"
  timeout 600 python bench.py --steps 5 --warmup 3 --frames $1 --res $2 --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/r02_call1_configs.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, {k:v['frac'] for k,v in d['kernels'].items()})"
done
