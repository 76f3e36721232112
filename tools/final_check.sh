#!/bin/bash
# Round-end rehearsal on one GPU: what the driver runs (tests, smoke, both bench arms) + the ncu launch list
# of the bench command itself.   gpurun --timeout 1500 -- 'bash tools/final_check.sh v10'
TAG=${1:-vX}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/r1_${TAG}_pytest_gpu.log 2>&1; tail -1 $O/r1_${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > $O/r1_${TAG}_bench_reference.json 2> $O/bench_ref.err
timeout 600 python bench.py > $O/r1_${TAG}_bench_bf16.json 2> $O/bench.err; cp $O/bench_layers_bf16.json $O/r1_${TAG}_layers_bf16.json
cat $O/r1_${TAG}_bench_bf16.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r1_${TAG}_ncu_launches_bench_cmd.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
tail -2 $O/ncu_bench.log | cut -c1-200
