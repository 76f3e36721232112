#!/bin/bash
# ncu evidence for the remaining kernels: single-CTA tcgen05 (upsamplers, resident k=3), conv_post, repack; fp32-mode launch list
TAG=${1:-vX}; O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none -k regex:"conv_tc_kernel|conv_post32|mel_to_operand" --launch-skip 0 --launch-count 14 \
  -f -o /tmp/prof_rest python tools/ncu_one_forward.py bf16 > $O/ncu_rest.log 2>&1
ncu -i /tmp/prof_rest.ncu-rep --page raw --csv > $O/r1_${TAG}_ncu_full_rest.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --launch-skip 79 --launch-count 79 --csv --log-file $O/r1_${TAG}_ncu_launches_fp32.csv python tools/ncu_one_forward.py fp32 > $O/ncu_launches_fp32.log 2>&1
ls -la $O/r1_${TAG}_ncu_full_rest.csv $O/r1_${TAG}_ncu_launches_fp32.csv
