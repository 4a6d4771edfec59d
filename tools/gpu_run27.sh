#!/bin/bash
mkdir -p gpurun_out
for t in kernels nets rcan swin; do timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x 2>&1 | tail -2; done
for c in rcan swinir_b16 edsr_m; do timeout 600 python tools/prof_step.py $c 3 2>&1 | grep -v Warn | tee gpurun_out/prof_step_$c.txt | head -16 | cut -c1-150; done
