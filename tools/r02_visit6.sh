#!/bin/bash
# round 2: BASELINE configs[4] sweep + configs[3] (single GPU form) + the 40 GiB single-GPU point (north-star target config)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== sweep"
timeout 1200 python tools/sweep.py --out gpurun_out/r02_sweep.json 2> gpurun_out/r02_sweep.err | tail -2
grep -c read_len gpurun_out/r02_sweep.err; tail -3 gpurun_out/r02_sweep.err | cut -c1-300
echo "== 40 GiB index, 10 M x 150 bp SE reads, one GPU"
timeout 1500 python bench.py --db-gib 40 --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/r02_bench_40gib.err | tail -1 > gpurun_out/r02_bench_40gib.json
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_40gib.json").read())
    print(round(d["value"]/1e6,2), "Mreads/s =", round(d["value"]*60/1e6,1), "M reads/min;", round(d["ms_per_step"],1), "ms; e2e", round(d["e2e"]["value"]/1e6,2), "frac", round(d["roofline"]["frac"],4))
    print({k: round(v,1) for k,v in d["stages_ms_per_step"].items()})
    print({k: v for k, v in d["config"].items() if k not in ("timing", "workload", "parallelism")})
except Exception as e:
    print("unparsable", e)
PY
tail -3 gpurun_out/r02_bench_40gib.err
