#!/bin/bash
# Host-side ceiling of `metabuli-b200 classify`: the real CLI against the CPU test double in null mode (tests/host/stub_backend.cpp,
# MBL_STUB_NULL=1: every batch "classifies" instantly), i.e. reader + batch pipeline + row formatter + writer alone.  No GPU needed.
#   usage: tools/host_ceiling.sh <db dir> <reads file> [<reads file 2>] [-- extra classify flags]
set -euo pipefail
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=$ROOT/tests/host/_build/cli_stub
[ -x $B/metabuli-b200 ] || { echo "run: python -m pytest tests/test_cli_cpu.py  (builds the double)"; exit 1; }
db=$1; shift
files=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do files+=("$1"); shift; done
[ $# -gt 0 ] && shift
mode=1; [ ${#files[@]} -eq 2 ] && mode=2
out=$(mktemp -d)
s=$(date +%s.%N)
MBL_STUB_NULL=1 MBL_STUB_DB_DIR=$db $B/metabuli-b200 classify --seq-mode $mode "$@" "${files[@]}" $db $out job | grep -E "Total read count|completed"
e=$(date +%s.%N)
n=$(($(wc -l < $out/job_classifications.tsv) - 1))
python3 -c "print('%d reads in %.2f s = %.2f M reads/s (host pipeline alone)' % ($n, $e - $s, $n / ($e - $s) / 1e6))"
rm -rf $out
