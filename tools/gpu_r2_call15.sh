#!/bin/bash
# Round 2, call 15: capped persistent grids for scatter / resolve so the two chunk streams co-run on every SM.
mkdir -p gpurun_out
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_sequence.py tests/test_gpu_full_sizes.py -x -q 2>&1 | tail -2
b() { timeout 600 python bench.py --steps 4 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; }
for cfg in "1024 1024" "1500 512" "75 256"; do set -- $cfg; export F=$1 R=$2
  for sc in 0 5 4 3; do for rc in 0 8 4 3 2; do
    echo "== frames=$F res=$R scatter_ctas=$sc resolve_ctas=$rc"; VPB200_SCATTER_CTAS=$sc VPB200_RESOLVE_CTAS=$rc b
  done; done
done
