#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_rcan.py -m gpu -q -x > gpurun_out/t_rcan.log 2>&1; echo "rcan exit=$?" >> gpurun_out/summary.txt
tail -15 gpurun_out/t_rcan.log >> gpurun_out/summary.txt
timeout 600 python tools/bench_all.py --steps 10 --only rcan > gpurun_out/bench_rcan.jsonl 2> gpurun_out/bench_rcan.err; echo "bench exit=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench_rcan.err >> gpurun_out/summary.txt
cat gpurun_out/bench_rcan.jsonl >> gpurun_out/summary.txt
timeout 600 python tools/prof_step.py rcan > gpurun_out/prof_rcan.txt 2>&1; head -30 gpurun_out/prof_rcan.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
