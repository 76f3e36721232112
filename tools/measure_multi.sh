#!/bin/bash
# Multi-GPU evidence run: gpurun --gpus 8 --timeout 1500 -- 'bash tools/measure_multi.sh v10'
TAG=${1:-vX}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 30 --warmup 3 > $O/r1_${TAG}_bench_bf16_n1.json 2> $O/bench_n1.err
  else timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 30 --warmup 3 --no-cpu-baseline > $O/r1_${TAG}_bench_bf16_n$n.json 2> $O/bench_n$n.err; fi
  tail -c 600 $O/r1_${TAG}_bench_bf16_n$n.json | head -c 0; python -c "import json; d=json.loads(open('$O/r1_${TAG}_bench_bf16_n$n.json').read().strip().splitlines()[-1]); print('N=$n', round(d['value']), 'audio-s/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value']), d['clocks'])"
done
timeout 400 $TR --nproc-per-node 8 --master-port 29541 tools/scaling_configs.py > $O/r1_${TAG}_scaling_configs_n8.jsonl 2> $O/scaling.err; cat $O/r1_${TAG}_scaling_configs_n8.jsonl
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/multi_gpu_check.py > $O/r1_${TAG}_multi_gpu_check_n8.json 2> $O/mgc.err; cat $O/r1_${TAG}_multi_gpu_check_n8.json
timeout 300 python -m pytest tests/test_gpu_e2e.py -q -m gpu -k multi_gpu 2>&1 | tail -2
