#!/bin/bash
# round-2 evidence: one `ncu --set full` capture per hot-path kernel at its BASELINE shape (tools/one_kernel.py), summarised
# into gpurun_out/r02_ncu_<name>.txt (copied to profiles/ by hand); each target first runs to exit 0 without ncu.
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; regex=${spec##*:}
  timeout 120 python tools/one_kernel.py $name --time > gpurun_out/r02_time_$name.txt 2>/dev/null || { echo "$name plain run failed"; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -f -o gpurun_out/r02_$name python tools/one_kernel.py $name > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/r02_$name.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1  python tools/one_kernel.py $name   (device time of back-to-back launches without ncu: $(cat gpurun_out/r02_time_$name.txt))" > gpurun_out/r02_ncu_$name.txt
  head -12 gpurun_out/r02_ncu_$name.txt
  [ -z "$KEEP_REP" ] && rm -f gpurun_out/r02_$name.ncu-rep   # (64 MiB limit on what travels back)
done
