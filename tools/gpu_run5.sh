#!/bin/bash
# launch list + full ncu capture of the two dominant kernels (eager launches so kernels are individually visible)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 520 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tapgemm_kernel -s 200 -c 2 -o gpurun_out/prof_tapgemm $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu tapgemm exit=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 100 -c 2 -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu wgrad exit=$?" >> gpurun_out/summary.txt
ls -la gpurun_out >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
