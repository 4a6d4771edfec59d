#!/bin/bash
# round-1 evidence run: full GPU test suite, smoke, both bench arms, ncu launch list + full-set captures
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit=$?" >> gpurun_out/summary.txt; tail -2 gpurun_out/t_all.log >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/summary.txt; tail -1 gpurun_out/smoke.log >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench.json >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit=$?" >> gpurun_out/summary.txt; cat gpurun_out/bench_ref.json >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2400 -c 1100 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu26.log 2>&1; echo "ncu list exit=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 10 -c 4 -o gpurun_out/prof_tapgemm_v3 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu26b.log 2>&1; echo "ncu full tapgemm exit=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad2 -s 10 -c 2 -o gpurun_out/prof_wgrad_v3 -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu26c.log 2>&1; echo "ncu full wgrad exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
