#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_n2.json >> gpurun_out/summary.txt; tail -8 gpurun_out/bench_n2.err >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/bench_ref.json >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
