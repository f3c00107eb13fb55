"""Oracle vs the reference binary on seeded synthetic databases (multi-species, LCA ties, Ns, ragged and
long reads, paired ends): the TSVs in tests/golden/synth were written by the reference itself
(tests/golden/gen_synth_golden.py).  Also checks that the synthetic generator is reproducible here."""
import gzip
import os

import pytest

import oracle
import synth_cases


@pytest.mark.parametrize("name", list(synth_cases.CASES) + list(synth_cases.CPU_CASES) + list(synth_cases.EDGE_CASES))
def test_oracle_matches_reference_on_synthetic(name, golden_dir, tmp_path):
    sdb, reads, seq_mode = synth_cases.build(name)
    want_fp = open(os.path.join(golden_dir, "synth", name + ".md5")).read().strip()
    assert synth_cases.fingerprint(sdb, reads) == want_fp, "synthetic inputs differ from the ones the golden TSV was made from"
    db_dir = str(tmp_path / "db")
    sdb.write(db_dir)
    mask, mask_prob = synth_cases.mask_flags(name)
    if mask:
        # --mask 1 cases: the reads the extractor sees are the masked ones (the product's host-side masking, itself pinned letter
        # by letter in test_host_mask.py); names and printed lengths stay the file's
        from test_host_mask import _mask
        reads = tuple(_mask(a, reads[k + 1], mask_prob) if k % 2 == 0 else a for k, a in enumerate(reads))
    q1 = str(tmp_path / "r1.fna")
    synth_cases.write_fasta(q1, reads[0], reads[1])
    q2 = None
    if seq_mode == 2:
        q2 = str(tmp_path / "r2.fna")
        synth_cases.write_fasta(q2, reads[2], reads[3])
    out = str(tmp_path / "o.tsv")
    oracle.classify_files(q1, q2, db_dir, seq_mode, out, threads=3, **synth_cases.oracle_flags(name))
    golden = gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
    assert open(out, "rb").read() == golden
    # <jobid>_report.tsv (Reporter::writeReportFile): clade counts walked from internal taxid 1, children by clade count (std::sort)
    want_report = gzip.open(os.path.join(golden_dir, "synth", name + ".report.gz"), "rb").read()
    assert open(out + ".report", "rb").read() == want_report
