#!/bin/bash
# GPU visit: the --mask 1 tests (device kernel K0 + host masker), a K0 timing, then the whole GPU suite.
tag=${1:-mask}
mkdir -p gpurun_out
echo "== pytest mask tests"
timeout 400 python -m pytest tests -m gpu -x -q -k "mask" > gpurun_out/r02_${tag}_pytest_mask.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r02_${tag}_pytest_mask.log
echo "== K0 timing"
timeout 300 python tools/mask_timing.py 2000000 150 > gpurun_out/r02_${tag}_timing.json 2> gpurun_out/r02_${tag}_timing.err; echo "rc=$?"; cat gpurun_out/r02_${tag}_timing.json; tail -3 gpurun_out/r02_${tag}_timing.err
echo "== pytest -m gpu (all)"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02_${tag}_pytest.log
