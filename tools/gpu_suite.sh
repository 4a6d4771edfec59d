#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for t in kernels nets rcan swin; do
timeout 900 python -m pytest tests/test_gpu_$t.py -m gpu -q -x > gpurun_out/t_$t.log 2>&1; echo "$t exit=$?" >> gpurun_out/summary.txt
tail -4 gpurun_out/t_$t.log >> gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench.err >> gpurun_out/summary.txt
cat gpurun_out/bench.json >> gpurun_out/summary.txt
timeout 1500 python tools/bench_all.py --steps 10 > gpurun_out/bench_all.jsonl 2> gpurun_out/bench_all.err; echo "bench_all exit=$?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench_all.err >> gpurun_out/summary.txt
python - >> gpurun_out/summary.txt <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','ms_per_step','patches_per_s','model_tflops','ms_per_tile','out_mpix_per_s')})
PY
cat gpurun_out/summary.txt
