#!/bin/bash
# round 2, visit 2: merge kernel v2 (CTA-wide balanced match stage, qinfo sorted with the value) — parity, speed, ncu
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-v2}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_${T}_pytest.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r02_${T}_pytest.log
summ() { python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(round(d['value']/1e6,2), 'Mreads/s', round(d['ms_per_step'],1), 'ms  frac', round(d['roofline']['frac'],4), {k: round(v,1) for k,v in d['stages_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']/1e6,2), d.get('cpu_baseline'), d.get('parity_sample'))
except Exception as e:
    print('unparsable', e)
"; }
echo "== bench default (with reference arm + parity sample)"
timeout 1200 python bench.py --steps 3 --warmup 3 2>gpurun_out/r02_${T}_bench_full.err | tail -1 | tee gpurun_out/r02_${T}_bench_full.json | summ
tail -3 gpurun_out/r02_${T}_bench_full.err
for cfg in "t256:MBL_MERGE_THREADS=256" "v1:MBL_MERGE_V1=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $tag ($envs)"
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02_${T}_bench_$tag.json | summ
done
KRE='regex:(merge_|extract_kernel|read_meta|score_|segment_kernel|match_|seq_bounds|taxcnt|RadixSort|DeviceScan|DeviceSelect|filter_|fg_len|read_len)'
echo "== ncu launch list (one step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 700 --csv --log-file gpurun_out/r02_${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_${T}_ncu_launch_run.log 2>&1
tail -1 gpurun_out/r02_${T}_ncu_launch_run.log | cut -c1-200
echo "== ncu full: merge kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 1 -c 1 -o gpurun_out/r02_${T}_merge_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_${T}_ncu_merge_run.log 2>&1
tail -1 gpurun_out/r02_${T}_ncu_merge_run.log | cut -c1-200
if [ "${2:-}" = "score" ]; then
echo "== ncu full: scoring kernels of one chunk"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:(score_fg_kernel|score_sp_kernel|score_kernel|match_gather_fix)' -c 4 -o gpurun_out/r02_${T}_score_prof -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_${T}_ncu_score_run.log 2>&1
tail -1 gpurun_out/r02_${T}_ncu_score_run.log | cut -c1-200
fi
ls -la gpurun_out | tail -8
