#!/bin/bash
# round 2, quick visit: parity + bench (no CPU arm) + ncu of the merge kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-v3}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_${T}_pytest.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02_${T}_pytest.log
summ() { python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(round(d['value']/1e6,2), 'Mreads/s', round(d['ms_per_step'],1), 'ms  frac', round(d['roofline']['frac'],4), {k: round(v,1) for k,v in d['stages_ms_per_step'].items()}, 'e2e', round(d['e2e']['value']/1e6,2), d.get('parity_sample'))
except Exception as e:
    print('unparsable', e)
"; }
echo "== bench default"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02_${T}_bench.err | tail -1 | tee gpurun_out/r02_${T}_bench.json | summ
tail -3 gpurun_out/r02_${T}_bench.err
shift
for cfg in "$@"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $tag ($envs)"
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02_${T}_bench_$tag.json | summ
done
if [ "${LAUNCHES:-0}" = "1" ]; then
echo "== ncu launch list (one step)"
KRE='regex:(merge_|extract_kernel|read_meta|score_|segment|match_|seq_|seg_|taxcnt|RadixSort|DeviceScan|DeviceSelect|filter_|fg_len|read_len|read_first)'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 700 --csv --log-file gpurun_out/r02_${T}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_${T}_ncu_launch_run.log 2>&1
tail -1 gpurun_out/r02_${T}_ncu_launch_run.log | cut -c1-120
fi
if [ "${SKIP_NCU:-0}" = "1" ]; then exit 0; fi
echo "== ncu full: merge kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_kernel -s 1 -c 1 -o gpurun_out/r02_${T}_merge_prof -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_${T}_ncu_merge_run.log 2>&1
tail -1 gpurun_out/r02_${T}_ncu_merge_run.log | cut -c1-200
