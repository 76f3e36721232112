#!/bin/bash
TAG=${1:-vX}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${2:-8}
timeout 500 $TR --nproc-per-node $N --master-port 29541 tools/scaling_configs.py > $O/r1_${TAG}_scaling_configs_n$N.jsonl 2> $O/scaling.err; cat $O/r1_${TAG}_scaling_configs_n$N.jsonl; tail -3 $O/scaling.err | cut -c1-300
timeout 300 $TR --nproc-per-node $N --master-port 29511 tools/multi_gpu_check.py > $O/r1_${TAG}_multi_gpu_check_n$N.json 2> $O/mgc.err; cat $O/r1_${TAG}_multi_gpu_check_n$N.json
