#!/bin/bash
mkdir -p gpurun_out
b() { timeout 600 python bench.py --steps 10 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; }
for cfg in "75 256" "1 256" "1024 1024"; do set -- $cfg; export F=$1 R=$2
  for c in none 100 50 25; do echo "== frames=$F res=$R carveout=$c"; if [ $c = none ]; then b; else VPB200_CARVEOUT=$c b; fi; done
done
