#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_rcan.py -m gpu -q -s > gpurun_out/t_rcan.log 2>&1; echo "rcan exit=$?" >> gpurun_out/summary.txt
grep -n "RCAN full\|passed\|failed\|Error\|assert" gpurun_out/t_rcan.log | head -20 >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "wgrad or colsum" > gpurun_out/t_kernels.log 2>&1; echo "kernels exit=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/t_kernels.log >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}, d['e2e']['value'], d['roofline']['achieved'])" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
