#!/bin/bash
mkdir -p gpurun_out
for t in kernels nets rcan swin; do timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x 2>&1 | tail -2; done
python tools/trace_tapgemm.py 2>&1 | grep -B1 -A5 "cycles relative" | cut -c1-700 | head -40
timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | grep -v wgrad | cut -c1-150
for c in swinir_b16; do timeout 600 python tools/prof_step.py $c 3 2>&1 | grep -v Warn | tee gpurun_out/prof_step_$c.txt | head -8 | cut -c1-150; done
