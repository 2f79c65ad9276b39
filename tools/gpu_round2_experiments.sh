#!/bin/bash
# First GPU call of round 2: run the three opt-in flavours that round 1 prepared without GPU time left, each under
# its own timeout (a hang must not cost the box), correctness first, then the A/B bench lines.
#   VPB200_BASIS_EPI=1     K1 epilogue through shared memory + cp.async.bulk stores
#   VPB200_VERTEX_SLOTS=1  bank-conflict-aware shared-memory slots of the fan vertex kernel
#   VPB200_HOST_PIPE=1     stage-parallel host-output pipeline (end-to-end path)
mkdir -p gpurun_out
export VPB200_TEST_EXPERIMENTAL=1
echo "== basis bulk epilogue: test"
VPB200_BASIS_EPI=1 timeout 120 python -m pytest tests/test_gpu_reconstruct.py -x -q -k "bulk_store or tensor_core" 2>&1 | tail -3
echo "== vertex slots: test"
timeout 180 python -m pytest tests/test_gpu_reconstruct.py -x -q -k "slot_flavour" 2>&1 | tail -3
echo "== host pipeline: test"
VPB200_HOST_PIPE=1 timeout 180 python -m pytest tests/test_gpu_sequence.py -x -q 2>&1 | tail -3
unset VPB200_TEST_EXPERIMENTAL
for cfg in "VPB200_NOP=1" "VPB200_BASIS_EPI=1" "VPB200_VERTEX_SLOTS=1" "VPB200_HOST_PIPE=1" "VPB200_BASIS_EPI=1 VPB200_VERTEX_SLOTS=1 VPB200_HOST_PIPE=1"; do
  echo "== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/r02_experiments.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
done
