#!/bin/bash
mkdir -p gpurun_out
for m in rexnet1_0x repvgg_a0; do
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02b_step_$m.csv python tools/dev_one_step.py $m 256 > gpurun_out/ncu_$m.log 2>&1
  tail -2 gpurun_out/ncu_$m.log
  wc -l gpurun_out/r02b_step_$m.csv
done
