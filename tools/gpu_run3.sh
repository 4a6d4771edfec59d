#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q > gpurun_out/t_nets.log 2>&1; echo "nets exit=$?" >> gpurun_out/summary.txt
tail -8 gpurun_out/t_nets.log >> gpurun_out/summary.txt
timeout 300 python tools/step_profile.py > gpurun_out/step_profile.json 2>gpurun_out/step_profile.err; cat gpurun_out/step_profile.json >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
