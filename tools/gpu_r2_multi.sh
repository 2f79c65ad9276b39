#!/bin/bash
# usage: tools/gpu_r2_multi.sh N [tests]  -- sharded12000 bench on N GPUs (torchrun), optionally the 2-GPU parity tests first
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$2" = "tests" ]; then
  echo "== 2-GPU parity tests"; timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -4
fi
for mode in ${MODES:-p2p}; do
  echo "== bench N=$N gather=$mode"
  pm=auto; g=$mode; case $mode in p2p-*) pm=${mode#p2p-}; g=p2p;; esac
  VPB200_GATHER=$g VPB200_PEER_MODE=$pm timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 2> gpurun_out/r02_multi_n${N}_$mode.err | tail -1 > gpurun_out/r02_multi_n${N}_$mode.json
  tail -3 gpurun_out/r02_multi_n${N}_$mode.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_multi_n${N}_$mode.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gather_verified','n1_same_config','gather','e2e')})
print('efficiency vs n1 same config:', d['value']/(d['n_gpus']*d['n1_same_config']['value']) if d.get('n1_same_config') else None)
print({k:v['ms'] for k,v in d.get('kernels',{}).items()})"
done
