#!/bin/bash
# compute-sanitizer over the small-model smoke path and the small GPU tests (racecheck covers the shared-memory
# double buffering of the vertex kernels, memcheck everything else)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python __graft_entry__.py smoke > gpurun_out/sanitize_racecheck.log 2>&1
tail -4 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py smoke > gpurun_out/sanitize_memcheck.log 2>&1
tail -4 gpurun_out/sanitize_memcheck.log
timeout 800 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_texture_normals.py tests/test_gpu_composite.py tests/test_gpu_shape_loss.py -x -q -k "not full and not large" > gpurun_out/sanitize_memcheck_tests.log 2>&1
tail -4 gpurun_out/sanitize_memcheck_tests.log
