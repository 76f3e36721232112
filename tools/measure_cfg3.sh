#!/bin/bash
# cfg-3 (64 x 2000 frames, utterance-sharded) at 1/2/4 GPUs:  gpurun --gpus 4 -- 'bash tools/measure_cfg3.sh v12'
TAG=${1:-vX}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4; do
  timeout 200 $TR --nproc-per-node $n --master-port $((29560 + n)) tools/scaling_configs.py --skip-long 2> $O/scaling_n$n.err | grep "^{" > $O/r1_${TAG}_scaling_configs_n$n.jsonl
  cat $O/r1_${TAG}_scaling_configs_n$n.jsonl | cut -c1-200
done
