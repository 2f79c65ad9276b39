#!/bin/bash
# usage: tools/gpu_multi.sh <ngpus> "ENV=.." ...   one torchrun bench per configuration ('-' = defaults)
n=$1; shift
port=29600
for cfg in "$@"; do
  [ "$cfg" = "-" ] && cfg="VPB200_NOP=1"
  port=$((port+1))
  echo "== N=$n $cfg"
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 2>&1 | tail -1 | tee -a gpurun_out/multi_n${n}.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), round(d['ms_per_step']*1e3,1))"
done
