#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python tools/bench_all.py --steps 10 > gpurun_out/bench_all.jsonl 2> gpurun_out/bench_all.err; echo "bench_all exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_all.jsonl >> gpurun_out/summary.txt; tail -5 gpurun_out/bench_all.err >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
