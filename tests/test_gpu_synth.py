"""GPU parity on seeded synthetic databases: end-to-end TSV against the reference binary's own output
(tests/golden/synth) and stage-by-stage against the oracle."""
import gzip
import os

import numpy as np
import pytest

import oracle
import synth_cases

pytestmark = pytest.mark.gpu
BLANK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _sorted_records(a):
    return np.sort(a, order=list(a.dtype.names))


@pytest.mark.parametrize("name", list(synth_cases.CASES))
def test_synthetic_case(name, golden_dir, tmp_path):
    from metabuli_b200 import Classifier, ClassifyOptions
    sdb, reads, seq_mode = synth_cases.build(name)
    want_fp = open(os.path.join(golden_dir, "synth", name + ".md5")).read().strip()
    assert synth_cases.fingerprint(sdb, reads) == want_fp
    clf = Classifier(None, ClassifyOptions(seq_mode=seq_mode), database=sdb.database)
    try:
        # end to end vs the reference's TSV
        res, pairs = clf.classify_batch(*reads)
        tsv = clf.format_tsv(synth_cases.names(reads[1].size - 1), res, pairs).encode()
        golden = gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
        assert tsv == golden
        # stages vs the oracle
        db_dir = str(tmp_path / "db")
        sdb.write(db_dir)
        odb = oracle.OracleDb(db_dir)
        gv, gq = clf.extract(*reads)
        ov, oq, cov1, cov2 = oracle.extract(*reads, kmer_format=sdb.database.params.kmer_format, syncmer=sdb.database.params.syncmer,
                                            smer_len=sdb.database.params.smer_len)
        assert gv.size == ov.size
        gm_ = gv != BLANK
        om_ = ((oq >> np.uint64(32)) & np.uint64(0x1FFFFFFF)) != 0
        g = np.stack([gv[gm_], gq[gm_]], 1)
        o = np.stack([ov[om_], oq[om_]], 1)
        assert np.array_equal(g[np.lexsort((g[:, 1], g[:, 0]))], o[np.lexsort((o[:, 1], o[:, 0]))])
        sv, sq = clf.sort_kmers(gv, gq)
        aa = sv >> np.uint64(24)
        assert np.all(aa[1:] >= aa[:-1])
        osv, osq = oracle.sort_kmers(ov, oq)
        om = odb.match(osv, osq)
        gm = clf.match(sv, sq)
        assert np.array_equal(_sorted_records(gm), _sorted_records(om))
        gs = clf.sort_matches(gm)
        os_ = oracle.sort_matches(om)
        assert np.array_equal(gs, os_)
        c2 = cov2 if seq_mode == 2 else None
        gres, gpairs = clf.score(gs, cov1, c2)
        ores, opairs = odb.score(os_, cov1, c2, seq_mode=seq_mode)
        for f in ("classification", "query_length", "taxcnt_len", "is_classified"):
            assert np.array_equal(gres[f], ores[f]), f
        assert np.array_equal(gres["score"].view(np.uint32), ores["score"].view(np.uint32))
        assert np.array_equal(gpairs, opairs)
        odb.close()
    finally:
        clf.close()


def test_empty_and_tiny_batches():
    from metabuli_b200 import Classifier, ClassifyOptions
    sdb, reads, _ = synth_cases.build("ragged_se")
    clf = Classifier(None, ClassifyOptions(seq_mode=1), database=sdb.database)
    try:
        res, pairs = clf.classify_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
        assert res.size == 0 and pairs.shape[0] == 0
        # a single read too short for any k-mer, and one read of Ns only
        b = np.frombuffer(b"ACGTACGTACGTACGTACGTACG" + b"N" * 150, dtype=np.uint8).copy()
        o = np.array([0, 23, 173], dtype=np.uint64)
        res, pairs = clf.classify_batch(b, o)
        assert list(res["is_classified"]) == [0, 0] and list(res["query_length"]) == [21, 147]
    finally:
        clf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sort_bit", [24, 32, 40])
def test_query_sort_granularity(sort_bit, golden_dir, monkeypatch):
    """The merge only needs the queries grouped per tile: every supported coarseness of the K2 sort (MBL_SORT_BIT, normally
    chosen at load time) must give the reference's TSV."""
    from metabuli_b200 import Classifier, ClassifyOptions
    monkeypatch.setenv("MBL_SORT_BIT", str(sort_bit))
    for name in ("multi_se", "format1_pe"):
        sdb, reads, seq_mode = synth_cases.build(name)
        clf = Classifier(None, ClassifyOptions(seq_mode=seq_mode), database=sdb.database)
        try:
            res, pairs = clf.classify_batch(*reads)
            tsv = clf.format_tsv(synth_cases.names(reads[1].size - 1), res, pairs).encode()
            assert tsv == gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
        finally:
            clf.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode,parts", [("multi_se", 1, 2), ("multi_pe", 1, 2), ("ragged_se", 1, 2), ("multi_se", 2, 2), ("multi_pe", 2, 3),
                                             ("ragged_se", 2, 5), ("format1_pe", 2, 4)])
def test_two_lane_pipeline(name, mode, parts, golden_dir, monkeypatch):
    """Large batches are cut into sub-batches that run on two pipeline lanes (two host threads, two streams, shared index):
    mode 1 = the lanes take alternate sub-batches whole, mode 2 = staggered (the back half K4-K5 of sub-batch k overlaps the
    front half K1-K3 of sub-batch k+1).  Forcing those paths on a small batch must still give the reference's TSV, including
    the per-read (taxid, count) lists."""
    from metabuli_b200 import Classifier, ClassifyOptions
    monkeypatch.setenv("MBL_PIPELINE", str(mode))      # off by default (profiles/r01_v10_config_sweep.md)
    monkeypatch.setenv("MBL_PIPELINE_PARTS", str(parts))
    monkeypatch.setenv("MBL_PIPELINE_MIN_READS", "64")
    sdb, reads, seq_mode = synth_cases.build(name)
    clf = Classifier(None, ClassifyOptions(seq_mode=seq_mode), database=sdb.database)
    try:
        for _ in range(2):          # second call reuses both workspaces
            res, pairs = clf.classify_batch(*reads)
            assert clf.stats()["sub_batches"] == parts
            tsv = clf.format_tsv(synth_cases.names(reads[1].size - 1), res, pairs).encode()
            assert tsv == gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
    finally:
        clf.close()


def test_streaming_prefetch_equals_batch_calls():
    """classify_stream (next batch uploaded while the current one is classified) == one classify_batch per batch."""
    from metabuli_b200 import Classifier, ClassifyOptions, multigpu
    sdb, reads, seq_mode = synth_cases.build("multi_pe")
    n = reads[1].size - 1
    cuts = [0, n // 5, n // 5, n // 2, n]                      # includes an empty batch
    batches = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        b1, o1 = multigpu.slice_batch(reads[0], reads[1], lo, hi)
        b2, o2 = multigpu.slice_batch(reads[2], reads[3], lo, hi)
        batches.append((b1, o1, b2, o2))
    clf = Classifier(None, ClassifyOptions(seq_mode=seq_mode), database=sdb.database)
    try:
        want = [clf.classify_batch(*b) for b in batches]
        got = list(clf.classify_stream(batches))
        assert len(got) == len(want)
        for (r1, p1), (r2, p2) in zip(got, want):
            assert np.array_equal(r1, r2) and np.array_equal(p1, p2)
        assert sum(int(r["is_classified"].sum()) for r, _ in got) > 500
    finally:
        clf.close()


def test_very_long_ragged_reads_vs_oracle():
    """BASELINE configs[3]/[4] shape in miniature: ragged long reads up to 50 kbp (seq-mode 3, denominator 1000), 8 % substitutions,
    some Ns — the CUDA path against the oracle on every read (positions beyond 2^15, thousands of windows per frame)."""
    from metabuli_b200 import Classifier, ClassifyOptions, synth
    sdb = synth.make_db(genera=3, species_per_genus=3, strains_per_species=2, codons=20000, seed=41)
    reads = synth.make_reads(sdb, 48, 50000, seed=42, sub_rate=0.08, n_rate=0.0005, length_jitter=47000)
    odb = oracle.OracleDb.from_synth(sdb)
    clf = Classifier(None, ClassifyOptions(seq_mode=3), database=sdb.database)
    try:
        res, pairs = clf.classify_batch(*reads)
        _, ores, nk, nm = odb.classify_arrays(*reads, seq_mode=3, threads=2)
        st = clf.stats()
        assert st["n_matches"] == nm and nm > 10000
        for f in ("classification", "query_length", "taxcnt_len", "is_classified"):
            assert np.array_equal(res[f], ores[f]), f
        assert np.array_equal(res["score"].view(np.uint32), ores["score"].view(np.uint32))
        assert int(res["is_classified"].sum()) > 20
    finally:
        clf.close()
        odb.close()


@pytest.mark.parametrize("name", list(synth_cases.CPU_CASES))
def test_cpu_pinned_cases_on_gpu(name, golden_dir):
    """The cases of synth_cases.CPU_CASES (asymmetric mates, non-default flags, accession-level databases, Skip_redundancy 0,
    --lineage, 50 kbp ragged reads) end to end on the CUDA path against the reference binary's TSV (green on a B200 since round 2)."""
    from metabuli_b200 import Classifier, ClassifyOptions
    sdb, reads, seq_mode = synth_cases.build(name)
    f = synth_cases.oracle_flags(name)
    opt = ClassifyOptions(seq_mode=seq_mode, min_score=f["min_score"], min_sp_score=f["min_sp_score"], tie_ratio=f["tie_ratio"],
                          min_cons_cnt=f["min_cons"], min_cons_cnt_euk=f["min_cons_euk"], accession_level=f["accession_level"])
    clf = Classifier(None, opt, database=sdb.database)
    try:
        res, pairs = clf.classify_batch(*reads)
        tsv = clf.format_tsv(synth_cases.names(reads[1].size - 1), res, pairs, lineage=bool(f["lineage"])).encode()
        assert tsv == gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
    finally:
        clf.close()


def test_batch_without_any_kmer():
    """Every read shorter than one k-mer window (and a batch of Ns): all rows come back unclassified with the covered length the
    reference prints (Q8), nothing is sorted or merged."""
    from metabuli_b200 import Classifier, ClassifyOptions
    sdb, _, _ = synth_cases.build("multi_se")
    clf = Classifier(None, ClassifyOptions(seq_mode=1), database=sdb.database)
    try:
        lens = np.array([1, 5, 23, 26, 20, 2], dtype=np.uint64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        bases = np.frombuffer(b"ACGT" * 64, dtype=np.uint8)[: int(off[-1])].copy()
        res, pairs = clf.classify_batch(bases, off)
        assert not res["is_classified"].any() and pairs.shape[0] == 0
        want = [orc_len for orc_len in (oracle.lib().orc_max_covered_length(int(x)) for x in lens)]
        assert res["query_length"].tolist() == want
        n_reads = 64
        off = (np.arange(n_reads + 1, dtype=np.uint64) * 150)
        res, pairs = clf.classify_batch(np.full(int(off[-1]), ord("N"), dtype=np.uint8), off)
        assert not res["is_classified"].any() and clf.stats()["n_query_kmers"] == 0
    finally:
        clf.close()


def _download_reads(clf, n_bytes, mate=1):
    """The resident (masked) letters of the last upload, for the device-masking checks."""
    import ctypes as C
    out = np.zeros(n_bytes, dtype=np.uint8)
    clf._check(clf.lib.mbl_download_reads(clf.ctx, mate, out.ctypes.data_as(C.c_void_p), n_bytes))
    return out


@pytest.mark.parametrize("name", ["mask_se", "mask_pe"])
def test_masked_queries(name, golden_dir, tmp_path):
    """--mask 1 (KmerExtractor.cpp:308-314) against the reference binary's TSV written with --mask 1 [--mask-prob p]: masked by the
    device kernel (K0, mbl_config.mask_mode), masked on the host (mbl_mask_reads), and through the C++ host in both modes."""
    import subprocess
    from metabuli_b200 import Classifier, ClassifyOptions
    sdb, reads, seq_mode = synth_cases.build(name)
    assert synth_cases.fingerprint(sdb, reads) == open(os.path.join(golden_dir, "synth", name + ".md5")).read().strip()
    mask, prob = synth_cases.mask_flags(name)
    golden = gzip.open(os.path.join(golden_dir, "synth", name + ".tsv.gz"), "rb").read()
    names = synth_cases.names(reads[1].size - 1)
    for on_host in (False, True):
        clf = Classifier(None, ClassifyOptions(seq_mode=seq_mode, mask=mask, mask_prob=prob, mask_on_host=on_host), database=sdb.database)
        try:
            before = [a.copy() for a in reads]
            res, pairs = clf.classify_batch(*reads)
            assert all(np.array_equal(a, b) for a, b in zip(before, reads)), "the caller's reads must not be modified"
            if not on_host:
                # the letters the device masked == the letters the host masks (itself pinned on the reference in test_host_mask.py)
                for mate in range(len(reads) // 2):
                    want = clf.mask_reads(reads[2 * mate], reads[2 * mate + 1])
                    got = _download_reads(clf, want.size, mate + 1)
                    assert np.array_equal(got, want), (on_host, mate, int((got != want).sum()))
                assert clf.stats()["ms_mask"] > 0
            assert clf.format_tsv(names, res, pairs).encode() == golden, on_host
        finally:
            clf.close()
    clf = Classifier(None, ClassifyOptions(seq_mode=seq_mode), database=sdb.database)
    try:
        res0, pairs0 = clf.classify_batch(*reads)
        assert clf.format_tsv(names, res0, pairs0).encode() != golden      # the flag matters on this case
    finally:
        clf.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "metabuli_b200", "_lib", "metabuli-b200")
    db_dir = str(tmp_path / "db")
    sdb.write(db_dir)
    files = [str(tmp_path / "r1.fna")]
    synth_cases.write_fasta(files[0], reads[0], reads[1])
    if seq_mode == 2:
        files.append(str(tmp_path / "r2.fna"))
        synth_cases.write_fasta(files[1], reads[2], reads[3])
    for extra in ([], ["--mask-host", "1"]):
        cmd = [exe, "classify", "--seq-mode", str(seq_mode), "--threads", "4", "--batch-reads", "700", "--mask", "1", "--mask-prob", str(prob)] + extra
        r = subprocess.run(cmd + files + [db_dir, str(tmp_path), "job"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:]
        assert open(tmp_path / "job_classifications.tsv", "rb").read() == golden, extra
        assert open(tmp_path / "job_report.tsv", "rb").read() == gzip.open(os.path.join(golden_dir, "synth", name + ".report.gz"), "rb").read()


def test_device_masking_of_odd_and_long_reads(golden_dir):
    """K0 on the inputs of tests/golden/synth/mask_misc.masked.gz (written by the reference's own tantan objects): empty and
    one-letter reads, Ns, lower case, IUPAC, periods around the 50-offset limit, 1-20 kb reads (the code window moves, thousands of
    rescalings) at three thresholds; and a batch whose longest read shrinks the number of warps in flight."""
    from metabuli_b200 import Classifier, ClassifyOptions
    sdb, _, _ = synth_cases.build("multi_se")
    b, o = synth_cases.mask_misc_reads()
    want_all = gzip.open(os.path.join(golden_dir, "synth", "mask_misc.masked.gz"), "rb").read()
    n_line = int(o[-1]) + o.size - 1
    for i, prob in enumerate(synth_cases.MASK_MISC_PROBS):
        clf = Classifier(None, ClassifyOptions(seq_mode=3, mask=1, mask_prob=prob), database=sdb.database)
        try:
            clf.classify_batch(b, o)
            got = _download_reads(clf, b.size)
            lines = b"".join(bytes(got[int(o[k]):int(o[k + 1])]) + b"\n" for k in range(o.size - 1))
            assert lines == want_all[i * n_line:(i + 1) * n_line], prob
        finally:
            clf.close()
    # long ragged reads: device == host letters
    sdb2, reads, _ = synth_cases.build("ont_ragged")
    clf = Classifier(None, ClassifyOptions(seq_mode=3, mask=1), database=sdb2.database)
    try:
        clf.classify_batch(*reads)
        assert np.array_equal(_download_reads(clf, reads[0].size), clf.mask_reads(reads[0], reads[1]))
    finally:
        clf.close()
