#!/bin/bash
# round 2, visit 1: experimental kernels of round 1 (never run on a GPU) — parity and speed against the default path
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu with the experimental gate open"
MBL_TEST_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_exp.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r02_pytest_exp.log
for cfg in "base:" "direct:MBL_MERGE_DIRECT=1" "warp:MBL_SCORE_WARP=1" "both:MBL_MERGE_DIRECT=1 MBL_SCORE_WARP=1"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  echo "== bench $tag ($envs)"
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02_v1_bench_$tag.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['stages_ms_per_step'], d['config'].get('classified_per_step'), d['config'].get('matches_per_step'))
except Exception as e:
    print('unparsable', e)
"
done
