#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3
python tools/trace_tapgemm.py 2>&1 | grep -A4 "edsr body" | cut -c1-900
python tools/trace_tapgemm.py 2>&1 | grep -A2 "== wgrad" | cut -c1-700
echo "## kpix128"; timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | grep "edsrL" | cut -c1-150
echo "## kpix64"; SRB_WG2_KPIX=64 timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | grep "edsrL.*wgrad" | cut -c1-150
