#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (our arm + reference arm), ncu launch list, full ncu captures of the merge and
# the extract kernel.  Everything lands in gpurun_out/; tools/summarize_ncu.py <tag> extract turns it into profiles/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; RC=$?; tail -5 gpurun_out/pytest.log
if [ $RC -ne 0 ]; then echo "== GPU TESTS FAILED (rc=$RC)"; tail -60 gpurun_out/pytest.log; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench full"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_full.json | cut -c1-400
if [ "${1:-}" != "noref" ]; then
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-300
fi
KRE='regex:(merge_|extract_kernel|read_meta|score_|segment_kernel|match_|seq_bounds|taxcnt|RadixSort|DeviceScan|DeviceSelect|filter_)'
echo "== ncu launch list (full workload, one step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
tail -1 gpurun_out/ncu_launch_run.log | cut -c1-200
echo "== ncu full: merge kernel and extract kernel (full workload)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 1 -c 1 -o gpurun_out/merge_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
tail -1 gpurun_out/ncu_full_run.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:extract_kernel -s 1 -c 1 -o gpurun_out/extract_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_extract_run.log 2>&1
tail -1 gpurun_out/ncu_extract_run.log | cut -c1-200
ls -la gpurun_out | tail -12
