#!/bin/bash
# final round-1 evidence: ncu launch list of bench.py (one replayed step) + full-set captures of the two GEMM kernels
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/plain29.log 2>&1; echo "plain exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2300 -c 1100 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu29.log 2>&1; echo "ncu list exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 10 -c 4 -o gpurun_out/prof_tapgemm_v4 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu29b.log 2>&1; echo "ncu full tapgemm exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad2 -s 10 -c 2 -o gpurun_out/prof_wgrad_v4 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu29c.log 2>&1; echo "ncu full wgrad exit=$?"
