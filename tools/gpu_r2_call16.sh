#!/bin/bash
# Round 2, call 16: triangles of a tile's run ordered by area (similar boxes per warp): parity + A/B.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512" "75 256" "12000 256"; do set -- $cfg; export F=$1 R=$2
  for s in 0 1; do echo "== frames=$F res=$R sort=$s"; VPB200_SORT_TRIS=$s b; done
done
