"""The reference's own regression checks for classify (util/Metabuli-regression/regression/run_classify_inclusion.sh and
run_classify_exclusion.sh), restated over the oracle (CPU): (1) FASTA and FASTQ inputs give the same <jobid>_report.tsv,
paired-end and single-end; (2) recall / precision read off the report reach the scripts' targets.  The GPU path is held to the
same files by tests/test_gpu_fixtures.py (its TSV and report are compared byte for byte with the reference binary's)."""
import os

import pytest

import oracle


def _run(fixtures_dir, tmp_path, db, mode, ext):
    q1 = os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_1.{ext}")
    q2 = os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_2.{ext}") if mode == "pe" else None
    out = str(tmp_path / f"{db}_{mode}_{ext.replace('.', '_')}.tsv")
    oracle.classify_files(q1, q2, os.path.join(fixtures_dir, f"db_{db}"), 2 if mode == "pe" else 1, out, threads=2)
    return open(out + ".report").read()


def _col(report, taxid, col):
    for ln in report.split("\n"):
        f = ln.split("\t")
        if len(f) >= 6 and f[4] == str(taxid):
            return float(f[col])
    raise AssertionError(f"taxid {taxid} not in the report")


@pytest.mark.parametrize("db", ["in", "ex"])
def test_fasta_and_fastq_give_the_same_report(db, fixtures_dir, tmp_path):
    for mode in ("pe", "se"):
        assert _run(fixtures_dir, tmp_path, db, mode, "fna.gz") == _run(fixtures_dir, tmp_path, db, mode, "fq.gz")


def _fmt(x):
    """bc scale=4 output as the scripts see it: truncated to four decimals, no leading zero (".9950")."""
    t = "%.4f" % (int(x * 10000) / 10000)
    return t[1:] if t.startswith("0.") else t


def _awk_verdict(actual: str, target: str) -> str:
    # awk -v actual="$ACTUAL" -v target="$TARGET" 'BEGIN { print (actual >= target) ? "GOOD" : "BAD" }': both are strings of four
    # numbers, so awk compares them as STRINGS — effectively the first number decides (the reference itself reaches only .9798
    # paired-end precision on the inclusion test at this commit, see the golden report)
    return "GOOD" if actual >= target else "BAD"


def test_inclusion_recall_and_precision(fixtures_dir, tmp_path):
    # run_classify_inclusion.sh: TARGET="12.1200 .9950 8.9800 .9955" (PE recall, PE precision, SE recall, SE precision)
    vals = []
    for mode in ("pe", "se"):
        rep = _run(fixtures_dir, tmp_path, "in", mode, "fna.gz")
        tp = _col(rep, 3000004, 1)
        classified = _col(rep, 2697049, 1) - _col(rep, 2697049, 2)
        vals += ["%.4f" % _col(rep, 3000004, 0), _fmt(tp / classified)]
    actual = " ".join(vals)
    assert _awk_verdict(actual, "12.1200 .9950 8.9800 .9955") == "GOOD", actual
    assert actual == "12.6200 .9798 9.8400 .9666"                 # = what the reference binary's own reports give (tests/golden/ref_tsv)


def test_exclusion_recall_and_precision(fixtures_dir, tmp_path):
    # run_classify_exclusion.sh: TARGET="76.0800 1.0000 71.2400 1.0000"
    vals = []
    for mode in ("pe", "se"):
        rep = _run(fixtures_dir, tmp_path, "ex", mode, "fna.gz")
        recall = _col(rep, 227984, 0)
        classified = float(rep.split("\n")[1].split("\t")[0])     # `head -n 2 | tail -n 1`: the line after the header
        vals += ["%.4f" % recall, _fmt(recall / classified)]
    actual = " ".join(vals)
    assert _awk_verdict(actual, "76.0800 1.0000 71.2400 1.0000") == "GOOD", actual
