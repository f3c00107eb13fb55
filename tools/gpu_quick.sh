#!/bin/bash
# Lean GPU-box visit: parity tests, smoke, bench (our arm), ncu launch list + full capture of the merge kernel.
# Everything lands in gpurun_out/.   usage: tools/gpu_quick.sh [tag]
set -u
TAG=${1:-cur}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; RC=$?; tail -5 gpurun_out/pytest_$TAG.log
if [ $RC -ne 0 ]; then echo "== GPU TESTS FAILED (rc=$RC)"; tail -60 gpurun_out/pytest_$TAG.log; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench full"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_full_$TAG.json
KRE='regex:(merge_|extract_kernel|read_meta|score_|segment_kernel|match_|seq_bounds|taxcnt|RadixSort|DeviceScan|DeviceSelect)'
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --db-gib 1 --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
tail -1 gpurun_out/ncu_launch_run.log | cut -c1-300
echo "== ncu full (merge kernel, FULL bench workload)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 1 -c 1 -o gpurun_out/merge_prof_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log | cut -c1-200
ls -la gpurun_out
