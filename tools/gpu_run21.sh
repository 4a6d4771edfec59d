#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -15
python tools/trace_tapgemm.py 2>&1 | grep -A5 "edsr body" | cut -c1-900
echo "## 2cta"; timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | grep "edsrL" | grep -v wgrad | cut -c1-150
echo "## 1cta"; SRB_TAPGEMM_1CTA=1 timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | grep "edsrL" | grep -v wgrad | cut -c1-150
