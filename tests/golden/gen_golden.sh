#!/bin/bash
# Regenerates the golden vectors of tests/golden/ from the REFERENCE ITSELF.  Container-only: needs
# /root/reference (read-only) and cmake; nothing here runs on the GPU box or in the test suite.
#
#  1. builds the unmodified reference out of tree in /tmp into oracle/_ref/metabuli (oracle/build_ref.sh, SURVEY.md §8c recipe),
#  2. runs `metabuli classify` on the regression fixtures (4 configs) -> tests/golden/ref_tsv/*.tsv.gz,
#  3. runs it on the seeded synthetic cases of tests/synth_cases.py (CASES and CPU_CASES, with the flags of FLAGS)
#     -> tests/golden/synth/<case>.tsv.gz, <case>.report.gz
#     (+ <case>.md5 = fingerprint of the generated inputs, so a test can tell "inputs differ" from "bug").
set -euo pipefail
REPO=$(cd "$(dirname "$0")/../.." && pwd)
bash $REPO/oracle/build_ref.sh
BIN=$REPO/oracle/_ref/metabuli
D=/root/reference/util/Metabuli-regression/data
OUT=$(mktemp -d)
mkdir -p $REPO/tests/golden/ref_tsv $REPO/tests/golden/synth
for db in in ex; do
  $BIN classify --seq-mode 1 $D/reads/ERR9594652_5000_1.fna $D/reference/$db $OUT ${db}_se --threads 1 --max-ram 6 >/dev/null
  $BIN classify $D/reads/ERR9594652_5000_1.fna $D/reads/ERR9594652_5000_2.fna $D/reference/$db $OUT ${db}_pe --threads 1 --max-ram 6 >/dev/null
  for m in se pe; do
    gzip -9 -n -c $OUT/${db}_${m}_classifications.tsv > $REPO/tests/golden/ref_tsv/${db}_${m}_classifications.tsv.gz
    gzip -9 -n -c $OUT/${db}_${m}_report.tsv > $REPO/tests/golden/ref_tsv/${db}_${m}_report.tsv.gz
  done
done
cd $REPO && python tests/golden/gen_synth_golden.py $BIN $OUT "$@"
rm -rf $OUT
