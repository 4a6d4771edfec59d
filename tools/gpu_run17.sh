#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in kernels nets rcan swin; do
timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x > gpurun_out/t_$t.log 2>&1; echo "$t exit=$?" >> gpurun_out/summary.txt
tail -2 gpurun_out/t_$t.log >> gpurun_out/summary.txt
done
echo "## kernels 2cta" >> gpurun_out/summary.txt
timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | cut -c1-200 > gpurun_out/kernels.jsonl; cat gpurun_out/kernels.jsonl >> gpurun_out/summary.txt
echo "## kernels 1cta wgrad" >> gpurun_out/summary.txt
SRB_WGRAD_1CTA=1 timeout 300 python tools/bench_kernels.py --iters 10 2>&1 | grep wgrad | cut -c1-200 >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
