#!/bin/bash
# Round 2, call 20: multi-block tcgen05 basis launches (one launch per group of >= 512 frames).
mkdir -p gpurun_out
echo "== basis tests"; timeout 600 python -m pytest tests/test_gpu_reconstruct.py -m gpu -x -q -k "basis" 2>&1 | tail -3
echo "== per-launch times"; timeout 300 python tools/diag_basis_blocks.py 2>&1 | tail -9
b() { timeout 600 python bench.py --steps 6 --warmup 3 --frames $F --res $R --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), {k:(round(v['ms']*1e3,1), v['frac'], v['launches']) for k,v in d['kernels'].items()})"; }
export F=4096 R=1024; for g in 96 512 1024; do echo "== 4096x1024 basis group $g"; VPB200_BASIS_FRAMES=$g b; done
export F=12000 R=256; for g in 96 512 1024; do echo "== 12000x256 basis group $g"; VPB200_BASIS_FRAMES=$g b; done
export F=1500 R=512; for g in 96 512; do echo "== 1500x512 basis group $g"; VPB200_BASIS_FRAMES=$g b; done
echo "== full sizes"; timeout 900 python -m pytest tests/test_gpu_full_sizes.py tests/test_gpu_sequence.py -m gpu -x -q 2>&1 | tail -3
