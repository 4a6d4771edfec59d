#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/one_shot_swin.py"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -s 14 -c 6 -o gpurun_out/prof_swin $CMD > gpurun_out/ncu_swin.log 2>&1
echo "exit=$?"; tail -3 gpurun_out/ncu_swin.log
