#!/bin/bash
# One-GPU evidence run for profiles/ (round 2): tests, bench arms, other configs, ncu launch list with DRAM bytes,
# full captures of the pair kernels.   gpurun --timeout 2400 -- 'bash tools/measure_round2.sh v7'
TAG=${1:-vX}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/r2_${TAG}_pytest_gpu.log; tail -1 $O/r2_${TAG}_pytest_gpu.log
timeout 900 python bench.py > $O/r2_${TAG}_bench_bf16_n1.json 2> $O/bench_bf16.err; cp $O/bench_layers_bf16.json $O/r2_${TAG}_layers_bf16.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_${TAG}_bench_reference.json 2> $O/bench_ref.err
timeout 600 python tools/config_perf.py > $O/r2_${TAG}_config_perf.jsonl 2> $O/config_perf.err
timeout 300 python tools/pair_modes.py > $O/r2_${TAG}_pair_kernel_ab.txt 2>&1
timeout 300 python tools/profile_config.py v2_narrow 16 800 bf16 > $O/r2_${TAG}_cfg4_layers_bf16.txt 2>&1
# launch list of one forward (61 launches; the first forward is warm-up): per-launch time and DRAM bytes
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip 61 --launch-count 61 --csv --log-file $O/r2_${TAG}_ncu_launches_bf16.csv python tools/ncu_one_forward.py bf16 > $O/ncu_launches.log 2>&1
# the same launch list without ncu's L2 flush between kernels: the DRAM traffic the step really has (profiles/r2_tile_order_l2.md)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
  --launch-skip 61 --launch-count 61 --csv --log-file $O/r2_${TAG}_ncu_launches_bf16_in_situ.csv python tools/ncu_one_forward.py bf16 > $O/ncu_launches2.log 2>&1
# the same pass over the bench command itself (the driver's profile convention): first 400 launches
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_${TAG}_ncu_launches_bench_cmd.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_bench_cmd.log 2>&1
# full-set captures: all 18 pair launches of one forward (both pair kernels) and the CTA-pair convs of stage 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair --launch-skip 18 --launch-count 18 \
  -f -o /tmp/prof_pair python tools/ncu_one_forward.py bf16 > $O/ncu_pair.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel --launch-skip 42 --launch-count 14 \
  -f -o /tmp/prof_tc2 python tools/ncu_one_forward.py bf16 > $O/ncu_tc2.log 2>&1
for r in tc2 pair; do
  ncu -i /tmp/prof_$r.ncu-rep --page raw --csv > $O/r2_${TAG}_ncu_full_${r}.csv 2>/dev/null
done
ls -la $O | tail -20
tail -c 2500 $O/r2_${TAG}_bench_bf16_n1.json
