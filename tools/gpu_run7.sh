#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in rcan swin; do
timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -s > gpurun_out/t_$t.log 2>&1; echo "$t exit=$?" >> gpurun_out/summary.txt
grep -n "full\|passed\|failed\|Error\|rel-L2\|err " gpurun_out/t_$t.log | head -30 >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
