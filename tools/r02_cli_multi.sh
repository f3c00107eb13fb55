#!/bin/bash
# the C++ host on N real devices: replicas and index-sharded, fixtures -> must equal the reference's TSVs
set -u
N=${1:-2}
mkdir -p gpurun_out /tmp/cli_out
F=tests/golden/fixtures
for mode in replica sharded; do
  for db in in ex; do
    extra=""; [ "$mode" = "sharded" ] && extra="--index-sharded 1"
    metabuli_b200/_lib/metabuli-b200 classify --seq-mode 2 --threads 8 --gpus $N $extra --batch-reads 1300 \
        $F/reads/ERR9594652_5000_1.fna.gz $F/reads/ERR9594652_5000_2.fna.gz $F/db_$db /tmp/cli_out ${mode}_$db > gpurun_out/r02_cli_${mode}_${db}_n$N.log 2>&1
    rc=$?
    got=$(md5sum < /tmp/cli_out/${mode}_${db}_classifications.tsv | cut -d' ' -f1)
    want=$(zcat tests/golden/ref_tsv/${db}_pe_classifications.tsv.gz | md5sum | cut -d' ' -f1)
    echo "$mode db_$db N=$N rc=$rc md5 $got $( [ "$got" = "$want" ] && echo EQUAL || echo DIFFERENT ) $(tail -1 gpurun_out/r02_cli_${mode}_${db}_n$N.log)"
  done
done
