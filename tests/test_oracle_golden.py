"""The oracle is pinned to the reference: on the regression fixtures of the reference's own test suite
(util/Metabuli-regression) its per-read TSV must equal, byte for byte, what the reference binary wrote
(tests/golden/ref_tsv, produced by tests/golden/gen_golden.sh) and the md5 / counts recorded in
SURVEY.md §8(c) / BASELINE.md §5."""
import gzip
import hashlib
import os

import pytest

import oracle

KNOWN = {  # (db, mode): (query k-mers, matches, classified, md5 of <job>_classifications.tsv)
    ("in", "se"): (1229412, 174845, 4909, "6cb4e3e881597696e2734d86473168b2"),
    ("in", "pe"): (2458568, 349237, 4919, "7d7e8ec13cdd1005ba597206a30fa85e"),
    ("ex", "se"): (1229412, 77545, 3562, "9d586a00129a4b56e01e84bc4492ee19"),
    ("ex", "pe"): (2458568, 154365, 3804, "25594a469cd73e19d7af73a4f3cfccd0"),
}


@pytest.mark.parametrize("db,mode", list(KNOWN))
def test_oracle_matches_reference_tsv(db, mode, fixtures_dir, golden_dir, tmp_path):
    q1 = os.path.join(fixtures_dir, "reads", "ERR9594652_5000_1.fna.gz")
    q2 = os.path.join(fixtures_dir, "reads", "ERR9594652_5000_2.fna.gz") if mode == "pe" else None
    out = str(tmp_path / "o.tsv")
    nk, nm = oracle.classify_files(q1, q2, os.path.join(fixtures_dir, f"db_{db}"), 2 if mode == "pe" else 1, out, threads=2)
    data = open(out, "rb").read()
    want_k, want_m, want_c, want_md5 = KNOWN[(db, mode)]
    assert (nk, nm) == (want_k, want_m)
    assert sum(1 for ln in data.split(b"\n")[1:] if ln.startswith(b"1\t")) == want_c
    assert hashlib.md5(data).hexdigest() == want_md5
    golden = gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_classifications.tsv.gz"), "rb").read()
    assert data == golden
    # <jobid>_report.tsv: walked from internal taxid 1, which is not the root in this taxonomy (Q11)
    report = gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_report.tsv.gz"), "rb").read()
    assert open(out + ".report", "rb").read() == report


def test_oracle_thread_invariance(fixtures_dir, tmp_path):
    q1 = os.path.join(fixtures_dir, "reads", "ERR9594652_5000_1.fna.gz")
    outs = []
    for t in (1, 5):
        out = str(tmp_path / f"o{t}.tsv")
        oracle.classify_files(q1, None, os.path.join(fixtures_dir, "db_in"), 1, out, threads=t)
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]
