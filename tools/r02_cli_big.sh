#!/bin/bash
# The C++ host against the real library on a GPU at a size where its own pipeline shows: 1 M read pairs (the regression fixture's
# reads x 200) vs the fixture database, .gz and plain input; md5 of the TSV against the reference's (tiled) TSV, wall time of the run.
set -u
mkdir -p gpurun_out /tmp/cliq
Q=tests/host/_build/gpu_quick
E=${MBL_CLI:-metabuli_b200/_lib/metabuli-b200}
if [ ! -f $Q/big_1.fna.gz ]; then   # inputs are not kept in the tree (113 MB): the fixture's reads x 200, the reference's TSV rows x 200
  mkdir -p $Q
  python3 - <<'PY'
import gzip, hashlib
F = 'tests/golden/fixtures/reads/'
Q = 'tests/host/_build/gpu_quick/'
for k in (1, 2):
    raw = gzip.open(F + 'ERR9594652_5000_%d.fna.gz' % k).read()
    with gzip.open(Q + 'big_%d.fna.gz' % k, 'wb', compresslevel=4) as f:
        for i in range(200):
            f.write(raw)
gold = gzip.open('tests/golden/ref_tsv/in_pe_classifications.tsv.gz').read()
nl = gold.index(b"\n") + 1
open(Q + 'big.want.md5', 'w').write(hashlib.md5(gold[:nl] + gold[nl:] * 200).hexdigest() + "\n")
PY
fi
want=$(cat $Q/big.want.md5)
nproc
for mode in gz plain; do
  if [ $mode = plain ]; then gzip -dc $Q/big_1.fna.gz > /tmp/cliq/b1.fna; gzip -dc $Q/big_2.fna.gz > /tmp/cliq/b2.fna; f1=/tmp/cliq/b1.fna; f2=/tmp/cliq/b2.fna; else f1=$Q/big_1.fna.gz; f2=$Q/big_2.fna.gz; fi
  for rep in 1 2; do
    s=$(date +%s%N)
    $E classify --seq-mode 2 $f1 $f2 tests/golden/fixtures/db_in /tmp/cliq big_$mode > gpurun_out/r02_big_$mode.log 2>&1; rc=$?
    e=$(date +%s%N)
    got=$(md5sum < /tmp/cliq/big_${mode}_classifications.tsv | cut -d' ' -f1)
    echo "$mode run$rep rc=$rc $( [ "$got" = "$want" ] && echo EQUAL || echo DIFFERENT ) wall $(( (e - s) / 1000000 )) ms; $(tail -1 gpurun_out/r02_big_$mode.log)"
  done
done
