#!/bin/bash
# Shortest possible check of the C++ host against the REAL library on a GPU (no Python): fixtures and the --mask 1 case, md5 of the
# TSV against the reference binary's.  Inputs of the mask case are prepared on the CPU side under tests/host/_build/gpu_quick.
set -u
mkdir -p gpurun_out /tmp/cliq
F=tests/golden/fixtures
Q=tests/host/_build/gpu_quick
E=metabuli_b200/_lib/metabuli-b200
run() {  # label want_file_cmd args...
  local label=$1 want=$2; shift 2
  $E classify "$@" /tmp/cliq $label > gpurun_out/r02_quick_$label.log 2>&1; local rc=$?
  local got=$(md5sum < /tmp/cliq/${label}_classifications.tsv | cut -d' ' -f1)
  echo "$label rc=$rc $( [ "$got" = "$want" ] && echo EQUAL || echo DIFFERENT ) $(tail -1 gpurun_out/r02_quick_$label.log)"
}
w_in_pe=$(zcat tests/golden/ref_tsv/in_pe_classifications.tsv.gz | md5sum | cut -d' ' -f1)
w_ex_se=$(zcat tests/golden/ref_tsv/ex_se_classifications.tsv.gz | md5sum | cut -d' ' -f1)
w_mask=$(md5sum < $Q/mask_se.want.tsv | cut -d' ' -f1)
run in_pe $w_in_pe --seq-mode 2 --threads 8 $F/reads/ERR9594652_5000_1.fna.gz $F/reads/ERR9594652_5000_2.fna.gz $F/db_in
run ex_se $w_ex_se --seq-mode 1 --threads 8 --batch-reads 1777 $F/reads/ERR9594652_5000_1.fq.gz $F/db_ex
run mask_dev $w_mask --seq-mode 1 --threads 8 --mask 1 $Q/mask_se.fna.gz $Q/db_mask_se
run mask_host $w_mask --seq-mode 1 --threads 8 --mask 1 --mask-host 1 --batch-reads 700 $Q/mask_se.fna $Q/db_mask_se
