#!/bin/bash
# usage: tools/sweep.sh "VAR=a VAR=b ..."   (each entry one env assignment; prints per-kernel ms)
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
