#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain19.log 2>&1; echo "plain exit=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu19.log 2>&1; echo "ncu list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 10 -c 3 -o gpurun_out/prof_tapgemm_v2 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu19b.log 2>&1; echo "ncu full tapgemm exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad -s 10 -c 2 -o gpurun_out/prof_wgrad_v2 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu19c.log 2>&1; echo "ncu full wgrad exit=$?"
wc -l gpurun_out/launches_r01c.csv; ls -la gpurun_out/*.ncu-rep
