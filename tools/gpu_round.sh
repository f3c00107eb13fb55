#!/bin/bash
# One GPU-box visit: parity tests, smoke, merge tuning, bench (our arm + reference arm), ncu launch list + full
# capture of the merge kernel.  Everything lands in gpurun_out/.   usage: tools/gpu_round.sh [quick]
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; RC=$?; tail -15 gpurun_out/pytest.log
if [ $RC -ne 0 ]; then echo "== GPU TESTS FAILED (rc=$RC): stopping here"; tail -60 gpurun_out/pytest.log; exit 1; fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
if [ "${1:-}" = "tune" ]; then echo "== tune only"; timeout 600 python tools/tune_merge.py --cells ${TUNE_CELLS:-2} --env "${TUNE_ENV:-MBL_DYN_CHUNKS=0,1}" 2>&1 | tail -8 | cut -c1-400; exit 0; fi
echo "== tune"; timeout 600 python tools/tune_merge.py --cells 1,2,4 2>&1 | tail -8 | tee gpurun_out/tune.jsonl
BEST=$(tail -1 gpurun_out/tune.jsonl | python -c "import json,sys; print(json.loads(sys.stdin.read()).get('best_tile_cells',4))" 2>/dev/null || echo 4)
export MBL_TILE_CELLS=${MBL_TILE_CELLS_FORCE:-$BEST}
echo "== using MBL_TILE_CELLS=$MBL_TILE_CELLS"
echo "== bench full"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_full.json
if [ "${1:-}" != "quick" ]; then
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_reference.json
fi
KRE='regex:(merge_|extract_kernel|read_meta|score_|segment_kernel|match_|seq_bounds|taxcnt|RadixSort|DeviceScan|DeviceSelect)'
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --db-gib 1 --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1
tail -1 gpurun_out/ncu_launch_run.log | cut -c1-300
echo "== ncu full (merge kernel, FULL bench workload; score kernel, reduced workload)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 1 -c 1 -o gpurun_out/merge_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_ -s 4 -c 4 -o gpurun_out/score_prof -f \
    python bench.py --db-gib 1 --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_score_run.log 2>&1
tail -1 gpurun_out/ncu_score_run.log | cut -c1-200
ls -la gpurun_out
