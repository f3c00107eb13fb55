#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_v7_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_v7_pytest.log
bash tools/r02_visit6.sh
