"""GPU parity on the reference's regression fixtures: every stage of the CUDA path, called through the
C-ABI, against the CPU oracle on the same inputs, and the end-to-end TSV against the reference binary's
own output (tests/golden/ref_tsv).  Bit-exact everywhere (integer work; float32 scores compared as bits)."""
import gzip
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

BLANK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _reads(fixtures_dir, mode):
    from metabuli_b200 import read_fastx
    n1, b1, o1 = read_fastx(os.path.join(fixtures_dir, "reads", "ERR9594652_5000_1.fna.gz"))
    if mode == "pe":
        _, b2, o2 = read_fastx(os.path.join(fixtures_dir, "reads", "ERR9594652_5000_2.fna.gz"))
        return n1, b1, o1, b2, o2
    return n1, b1, o1, None, None


def _sorted_records(a):
    return np.sort(a, order=list(a.dtype.names))


@pytest.fixture(scope="module", params=[("in", "se"), ("in", "pe"), ("ex", "se"), ("ex", "pe")], ids=lambda p: f"{p[0]}-{p[1]}")
def case(request, fixtures_dir):
    from metabuli_b200 import Classifier, ClassifyOptions
    db, mode = request.param
    db_dir = os.path.join(fixtures_dir, f"db_{db}")
    clf = Classifier(db_dir, ClassifyOptions(seq_mode=2 if mode == "pe" else 1))
    odb = oracle.OracleDb(db_dir)
    yield db, mode, clf, odb, _reads(fixtures_dir, mode)
    clf.close()
    odb.close()


def test_stage_parity(case):
    db, mode, clf, odb, (names, b1, o1, b2, o2) = case
    # --- K1 extract: same multiset of (value, qinfo), same number of blank slots
    gv, gq = clf.extract(b1, o1, b2, o2)
    ov, oq, cov1, cov2 = oracle.extract(b1, o1, b2, o2, kmer_format=odb.kmer_format)
    assert gv.size == ov.size
    gmask = gv != BLANK
    omask = (oq >> np.uint64(32)) & np.uint64(0x1FFFFFFF) != 0
    assert int(gmask.sum()) == int(omask.sum())
    g = np.stack([gv[gmask], gq[gmask]], 1)
    o = np.stack([ov[omask], oq[omask]], 1)
    g = g[np.lexsort((g[:, 1], g[:, 0]))]
    o = o[np.lexsort((o[:, 1], o[:, 0]))]
    assert np.array_equal(g, o)
    # --- K2 sort: amino-acid parts ascending, content preserved, blanks last
    sv, sq = clf.sort_kmers(gv, gq)
    aa = sv >> np.uint64(24)
    assert np.all(aa[1:] >= aa[:-1])
    s = np.stack([sv[sv != BLANK], sq[sv != BLANK]], 1)
    s = s[np.lexsort((s[:, 1], s[:, 0]))]
    assert np.array_equal(s, o)
    # --- K3 merge: same set of Match records as the reference's linear merge
    osv, osq = oracle.sort_kmers(ov, oq)
    om = odb.match(osv, osq)
    gm = clf.match(sv, sq)
    assert gm.size == om.size
    assert np.array_equal(_sorted_records(gm), _sorted_records(om))
    # --- K4 match sort: the reference order is total, so the arrays must be identical
    gs = clf.sort_matches(gm)
    os_ = oracle.sort_matches(om)
    assert np.array_equal(gs, os_)
    # --- K5 scoring
    seq_mode = 2 if mode == "pe" else 1
    gres, gpairs = clf.score(gs, cov1, cov2 if mode == "pe" else None)
    ores, opairs = odb.score(os_, cov1, cov2 if mode == "pe" else None, seq_mode=seq_mode)
    for f in ("classification", "hamming", "query_length", "taxcnt_len", "is_classified"):
        assert np.array_equal(gres[f], ores[f]), f
    assert np.array_equal(gres["score"].view(np.uint32), ores["score"].view(np.uint32))
    assert np.array_equal(gpairs, opairs)


def test_end_to_end_tsv_matches_reference(case, golden_dir):
    db, mode, clf, odb, (names, b1, o1, b2, o2) = case
    res, pairs = clf.classify_batch(b1, o1, b2, o2)
    tsv = clf.format_tsv(names, res, pairs).encode()
    golden = gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_classifications.tsv.gz"), "rb").read()
    assert tsv == golden
    st = clf.stats()
    known = {("in", "se"): (1229412, 174845), ("in", "pe"): (2458568, 349237), ("ex", "se"): (1229412, 77545), ("ex", "pe"): (2458568, 154365)}
    assert (st["n_query_kmers"], st["n_matches"]) == known[(db, mode)]
    assert st["kernel_launches"] > 0


@pytest.mark.parametrize("db,mode,batch,ext", [("in", "pe", 0, "fna.gz"), ("in", "se", 1500, "fna.gz"), ("ex", "pe", 777, "fna.gz"),
                                               ("ex", "se", 0, "fq.gz")])
def test_cpp_host_cli_writes_the_reference_tsv(db, mode, batch, ext, fixtures_dir, golden_dir, tmp_path):
    """The C++ host (`metabuli-b200 classify`, same command line as the reference): parallel FASTA reader, streamed batches
    (next batch uploads while the current one is classified), parallel TSV writer -> the reference binary's TSV, byte for byte."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "metabuli_b200", "_lib", "metabuli-b200")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    reads = [os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_{k}.{ext}") for k in ((1, 2) if mode == "pe" else (1,))]
    cmd = [exe, "classify", "--seq-mode", "2" if mode == "pe" else "1", "--threads", "4"]
    if batch:
        cmd += ["--batch-reads", str(batch)]
    cmd += reads + [os.path.join(fixtures_dir, f"db_{db}"), str(tmp_path), "job"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    got = open(tmp_path / "job_classifications.tsv", "rb").read()
    golden = gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_classifications.tsv.gz"), "rb").read()
    assert got == golden
    assert "Total read count : 5000" in r.stdout
    # the Kraken-style report (Reporter::writeReportFile)
    report = gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_report.tsv.gz"), "rb").read()
    assert open(tmp_path / "job_report.tsv", "rb").read() == report


@pytest.mark.parametrize("devices,batch", [("0,0", 400), ("0,0,0", 333)])
def test_cpp_host_multi_replica_keeps_the_read_order(devices, batch, fixtures_dir, golden_dir, tmp_path):
    """`--gpus N` / `--devices a,b,...`: one replica of the index per device, the batches of the input go to whichever device is
    free and the writer puts the rows back in input order.  Several contexts on device 0 exercise that host logic on a one-GPU
    box; many small batches make the devices finish out of order."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "metabuli_b200", "_lib", "metabuli-b200")
    reads = [os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_{k}.fna.gz") for k in (1, 2)]
    cmd = [exe, "classify", "--seq-mode", "2", "--threads", "3", "--devices", devices, "--batch-reads", str(batch)] + reads + \
          [os.path.join(fixtures_dir, "db_in"), str(tmp_path), "job"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    golden = gzip.open(os.path.join(golden_dir, "ref_tsv", "in_pe_classifications.tsv.gz"), "rb").read()
    assert open(tmp_path / "job_classifications.tsv", "rb").read() == golden
    report = gzip.open(os.path.join(golden_dir, "ref_tsv", "in_pe_report.tsv.gz"), "rb").read()
    assert open(tmp_path / "job_report.tsv", "rb").read() == report
    assert "Total read count : 5000" in r.stdout and f"on {len(devices.split(','))} GPU(s)" in r.stdout


@pytest.mark.parametrize("devices,batch,db,mode", [("0,0", 1700, "in", "pe"), ("0,0,0", 0, "ex", "se"), ("0,0,0,0", 640, "in", "se")])
def test_cpp_host_index_sharded(devices, batch, db, mode, fixtures_dir, golden_dir, tmp_path):
    """`--index-sharded 1`: the C++ host cuts the index into one value range per device (mbl_plan_shards), ORs the shards' presence
    filters, and runs every batch as one exchange round — extract + bucket, push metamers to the owning shard, match, push matches
    back to the read owner, score — with one thread per device in lock step.  Several contexts on device 0 exercise the whole
    protocol (peer pointers inside one process) on a one-GPU box; the TSV must be the reference's."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "metabuli_b200", "_lib", "metabuli-b200")
    reads = [os.path.join(fixtures_dir, "reads", f"ERR9594652_5000_{k}.fna.gz") for k in ((1, 2) if mode == "pe" else (1,))]
    cmd = [exe, "classify", "--seq-mode", "2" if mode == "pe" else "1", "--threads", "3", "--devices", devices, "--index-sharded", "1"]
    if batch:
        cmd += ["--batch-reads", str(batch)]
    cmd += reads + [os.path.join(fixtures_dir, f"db_{db}"), str(tmp_path), "job"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    golden = gzip.open(os.path.join(golden_dir, "ref_tsv", f"{db}_{mode}_classifications.tsv.gz"), "rb").read()
    assert open(tmp_path / "job_classifications.tsv", "rb").read() == golden
    assert "index sharded" in r.stdout and "Total read count : 5000" in r.stdout
