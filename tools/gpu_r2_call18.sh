#!/bin/bash
# Round 2, call 18: all GPU tests on the pruned build; vertex-kernel frames-per-CTA (one wave instead of two) A/B.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; }
export F=12000 R=256; for fpb in 0 93 62 31; do echo "== 12000x256 vertex fpb=$fpb"; VPB200_VERTEX_FPB=$fpb b; done
export F=1024 R=1024; for fpb in 0 10 19 38; do echo "== 1024x1024 vertex fpb=$fpb"; VPB200_VERTEX_FPB=$fpb b; done
export F=1500 R=512; for fpb in 0 14 27 54; do echo "== 1500x512 vertex fpb=$fpb"; VPB200_VERTEX_FPB=$fpb b; done
export F=75 R=256; for fpb in 0 10; do echo "== 75x256 vertex fpb=$fpb"; VPB200_VERTEX_FPB=$fpb b; done
