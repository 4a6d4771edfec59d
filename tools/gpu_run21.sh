#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in kernels nets; do
timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x > gpurun_out/t_$t.log 2>&1; echo "$t exit=$?" >> gpurun_out/summary.txt
tail -4 gpurun_out/t_$t.log >> gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt
timeout 200 python tools/prof_step.py edsr_l > gpurun_out/prof_step_edsr_l.txt 2>&1; grep "unpack_inline\|pack_weights\|ms/step" gpurun_out/prof_step_edsr_l.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
