#!/bin/bash
# First GPU bring-up: each test group in its own process (a trap in one kernel must not poison the rest).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, pytest -k expr
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$2" > gpurun_out/t_$1.log 2>&1
  echo "$1 exit=$?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/t_$1.log >> gpurun_out/summary.txt
}
rm -f gpurun_out/summary.txt
run remap "pixel_shuffle or window_remap or roll or layout"
run fwd "tapgemm_conv_fwd"
run epi "tapgemm_epilogues or pixel_shuffle_store"
run dgrad "tapgemm_dgrad"
run wgrad "wgrad"
timeout 600 python tools/bench_kernels.py --iters 10 > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err
echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
