#!/bin/bash
# One-GPU evidence run for profiles/: tests, bench arms, per-layer table, ncu launch list + full captures.
#   gpurun --timeout 2400 -- 'bash tools/measure_round.sh v10'
TAG=${1:-vX}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/r1_${TAG}_pytest_gpu.log 2>&1; tail -2 $O/r1_${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/r1_${TAG}_bench_bf16.json 2> $O/bench_bf16.err; cp $O/bench_layers_bf16.json $O/r1_${TAG}_layers_bf16.json 2>/dev/null
timeout 600 python bench.py --precision fp32 --no-cpu-baseline > $O/r1_${TAG}_bench_fp32.json 2> $O/bench_fp32.err; cp $O/bench_layers_fp32.json $O/r1_${TAG}_layers_fp32.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > $O/r1_${TAG}_bench_reference.json 2> $O/bench_ref.err
timeout 300 python tools/postnet_bench.py > $O/r1_${TAG}_postnet_bench.json 2> $O/postnet_bench.err
timeout 300 python tools/ragged_bench.py > $O/r1_${TAG}_ragged_bench.json 2> $O/ragged_bench.err
timeout 600 python tools/config_perf.py > $O/r1_${TAG}_config_perf.jsonl 2> $O/config_perf.err
# launch list of one forward (61 launches; the first forward is warm-up)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip 61 --launch-count 61 --csv --log-file $O/r1_${TAG}_ncu_launches_bf16.csv python tools/ncu_one_forward.py bf16 > $O/ncu_launches.log 2>&1
# full-set captures of the dominant kernels: CTA-pair conv (stage 1) and the fused ResBlock pair (stage 2/3)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel --launch-skip 12 --launch-count 14 \
  -f -o /tmp/prof_tc2 python tools/ncu_one_forward.py bf16 > $O/ncu_tc2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair_tc_kernel --launch-skip 0 --launch-count 18 \
  -f -o /tmp/prof_pair python tools/ncu_one_forward.py bf16 > $O/ncu_pair.log 2>&1
# the .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): export the raw page of every captured
# launch and the source page (per-instruction stalls, needs -lineinfo) of one launch of each kernel
for r in tc2 pair; do
  ncu -i /tmp/prof_$r.ncu-rep --page raw --csv > $O/r1_${TAG}_ncu_full_${r}.csv 2>/dev/null
done
ncu -i /tmp/prof_tc2.ncu-rep --page source --csv --launch-skip 12 --launch-count 1 2>/dev/null | gzip > $O/r1_${TAG}_ncu_source_tc2_s1k11c1.csv.gz
ncu -i /tmp/prof_pair.ncu-rep --page source --csv --launch-skip 6 --launch-count 1 2>/dev/null | gzip > $O/r1_${TAG}_ncu_source_pair_c64k11.csv.gz
ls -la $O | tail -30
cat $O/r1_${TAG}_bench_bf16.json
