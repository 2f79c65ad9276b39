#!/bin/bash
# Round 2, call 2: the fused vertex + raster kernel: parity first, then A/B bench lines at the BASELINE configs.
mkdir -p gpurun_out
echo "== fused parity tests"
timeout 600 python -m pytest tests/test_gpu_sequence.py -x -q 2>&1 | tail -15
echo "== slot flavour (fixed staging)"
VPB200_TEST_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_gpu_reconstruct.py -x -q -k "slot_flavour" 2>&1 | tail -3
echo "== pytest gpu (all)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "75 256" "1500 512" "4096 1024"; do
  set -- $cfg
  for fm in 0 1; do
    echo "== bench frames=$1 res=$2 separate=$fm"
    VPB200_BENCH_SEPARATE=$fm timeout 600 python bench.py --steps 5 --warmup 3 --frames $1 --res $2 --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/r02_call2.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
  done
done
for mb in 48 96; do
  echo "== 4096x1024 fused, chunk budget $mb MB"
  VPB200_CHUNK_MB=$mb timeout 600 python bench.py --steps 5 --warmup 3 --frames 4096 --res 1024 --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/r02_call2.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
