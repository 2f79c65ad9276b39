#!/bin/bash
# Round 2, call 6: two-pixel inline loop A/B, chunk-size (L2 residency of the z-buffer keys) x stream sweep.
mkdir -p gpurun_out
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_raster.py tests/test_gpu_full_sizes.py tests/test_gpu_reconstruct.py -x -q 2>&1 | tail -3
b() { timeout 600 python bench.py --steps 4 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, {k:v['launches'] for k,v in d['kernels'].items()})"; }
for cfg in "75 256" "1500 512" "1024 1024"; do
  set -- $cfg; export F=$1 R=$2
  for p in 0 1; do echo "== frames=$F res=$R pairs=$p"; VPB200_SCATTER_PAIRS=$p b; done
done
for cfg in "750 512" "512 1024"; do
  set -- $cfg; export F=$1 R=$2
  for dual in 1 0; do for mb in 24 40 64 96 128 192 384; do
    echo "== frames=$F res=$R dual=$dual chunk_mb=$mb"; VPB200_DUAL=$dual VPB200_CHUNK_MB=$mb b
  done; done
done
