#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in rcan nets; do
timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x > gpurun_out/t_$t.log 2>&1; echo "$t exit=$?" >> gpurun_out/summary.txt
tail -4 gpurun_out/t_$t.log >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
