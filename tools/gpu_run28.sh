#!/bin/bash
mkdir -p gpurun_out
echo "== tests without PDL"; timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
echo "== tests with PDL"; SRB_PDL=1 timeout 1400 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
for i in 1 2; do
echo "== bench no PDL"; python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
echo "== bench PDL"; SRB_PDL=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
echo "== bench_all PDL"; SRB_PDL=1 timeout 900 python tools/bench_all.py --steps 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('config','ms_per_step','patches_per_s','ms_per_tile','out_mpix_per_s')})"
