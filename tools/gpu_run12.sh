#!/bin/bash
mkdir -p gpurun_out
for c in swinir_b16 edsr_l rcan; do timeout 600 python tools/trace_step.py $c > gpurun_out/trace_$c.txt 2>&1; cat gpurun_out/trace_$c.txt | head -24; done
