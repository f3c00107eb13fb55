#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-k4}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_${T}_pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02_${T}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02_${T}_bench.err | tail -1 | tee gpurun_out/r02_${T}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(round(d['value']/1e6,2), 'Mreads/s', round(d['ms_per_step'],1), 'ms', {k: round(v,1) for k,v in d['stages_ms_per_step'].items() if v})
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_order_kernel -s 5 -c 2 -o gpurun_out/r02_${T}_order_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_${T}_ncu_order_run.log 2>&1
tail -1 gpurun_out/r02_${T}_ncu_order_run.log | cut -c1-100
