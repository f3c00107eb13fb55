#!/bin/bash
# round 2: bench.py at N GPUs (replica leg + index-sharded configs[2] leg in one line)
set -u
N=${1:-2}; STEPS=${2:-3}; WARM=${3:-2}; shift 3 || true
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1; nvidia-smi nvlink -gt d -i 0 > gpurun_out/r02_nvlink_raw_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup $WARM "$@" \
   > gpurun_out/r02_bench_n$N.out 2> gpurun_out/r02_bench_n$N.err
echo "rc=$?"
tail -1 gpurun_out/r02_bench_n$N.out > gpurun_out/r02_bench_n$N.json
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n$N.json").read())
    print("replica:", round(d["value"]/1e6,2), "Mreads/s", round(d["ms_per_step"],1), "ms", "e2e", round(d["e2e"]["value"]/1e6,2))
    s = d.get("sharded") or {}
    if "error" in s: print("SHARDED ERROR:", s["error"]); print(s.get("trace"))
    else:
        print("sharded:", round(s["value"]/1e6,2), "Mpairs/s", round(s["ms_per_step"],1), "ms", s["workload"])
        print(" phases", s["phases_ms_per_step_rank0"]); print(" a2a", s["a2a_rank0"]); print(" nvlink", s["nvlink_rank0"]); print(" stages", s["stages_ms_per_step_rank0"])
        print(" gen", s["db_gen_s"], "load", s["db_load_s"], "classified", s["classified_pairs_per_step"], "merge", s["merge_roofline_rank0"])
except Exception as e:
    print("unparsable:", e)
PY
tail -5 gpurun_out/r02_bench_n$N.err
