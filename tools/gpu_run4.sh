#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -s > gpurun_out/t_nets.log 2>&1; echo "nets exit=$?" >> gpurun_out/summary.txt
grep -n "grad rel-L2\|passed\|failed\|Error" gpurun_out/t_nets.log | head -20 >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/t_kernels.log 2>&1; echo "kernels exit=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/t_kernels.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt; tail -5 gpurun_out/bench.err >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "bench eager exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_eager.json >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
