#!/bin/bash
# Round 2, call 17: per-triangle colours indexed by original id (one gather in the resolve pass) A/B; 75-frame chunking.
mkdir -p gpurun_out
echo "== parity (colour by orig)"; VPB200_COLOR_BY_ORIG=1 timeout 600 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -x -q 2>&1 | tail -2
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512" "75 256" "12000 256"; do set -- $cfg; export F=$1 R=$2
  for s in 0 1; do echo "== frames=$F res=$R color_by_orig=$s"; VPB200_COLOR_BY_ORIG=$s b; done
done
export F=75 R=256
echo "== 75x256 one chunk"; VPB200_CHUNK_FRAMES=75 b
echo "== 75x256 three chunks"; VPB200_CHUNK_FRAMES=25 b
for fpb in 4 7 10 19; do echo "== 75x256 vertex fpb=$fpb"; VPB200_VERTEX_FPB=$fpb b; done
for fpb in 1 2 8; do echo "== 75x256 scatter fpb=$fpb"; VPB200_SCATTER_FPB=$fpb b; done
