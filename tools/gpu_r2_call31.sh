#!/bin/bash
# Round 2, call 31: chunk budget re-sweep on the final kernels; one stream against two.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r02_call31.log) 2>&1
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['launches']) for k,v in d['kernels'].items()})"; }
for cfg in "2048 1024" "1500 512" "3000 256"; do set -- $cfg; export F=$1 R=$2
  for mb in 160 240 320 480 720; do echo "== ${F}x${R} chunk_mb=$mb"; VPB200_CHUNK_MB=$mb b; done
  echo "== ${F}x${R} one stream"; VPB200_DUAL=0 b
done
