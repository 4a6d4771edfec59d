#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -x > gpurun_out/t_nets.log 2>&1; echo "nets exit=$?" >> gpurun_out/summary.txt
tail -15 gpurun_out/t_nets.log >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/smoke.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt; tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
