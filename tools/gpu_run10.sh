#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for cfg in swinir_b16 rcan; do
CMD="python tools/profile_step.py $cfg 2"
timeout 600 $CMD > gpurun_out/plain_$cfg.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$cfg.csv $CMD > gpurun_out/ncu_$cfg.log 2>&1
echo "$cfg ncu exit=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
