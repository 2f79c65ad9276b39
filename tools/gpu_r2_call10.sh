#!/bin/bash
# Round 2, call 10: GEMV-path check, K1 timeline of CTA 0, single-frame numbers.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== basis timing (all modes)"; timeout 200 python tools/diag_basis_time.py 2>&1 | grep -E "write\+read-flush|warm" 
echo "== K1 trace"; timeout 120 python tools/diag_basis_trace.py
echo "== single / grid"
for c in single grid; do timeout 300 python bench.py --steps 10 --warmup 3 --config $c --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; done
