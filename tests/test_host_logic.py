"""Host logic without a GPU: the product's scoring core (metabuli_b200/csrc/score_core.cuh, compiled for
the host by nvcc) against the oracle, and its libstdc++-order sort replay against std::sort (quirk Q4)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
import synth_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "host_logic.cu")
OUT = os.path.join(ROOT, "tests", "host", "_build", "libhost_logic.so")


@pytest.fixture(scope="module")
def hl():
    deps = [SRC, os.path.join(ROOT, "metabuli_b200", "csrc", "score_core.cuh"), os.path.join(ROOT, "metabuli_b200", "csrc", "kernels.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "--extended-lambda", "-fmad=false",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC])
    return C.CDLL(OUT)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def test_sort_replays_std_sort(hl):
    rng = np.random.default_rng(5)
    hl.ht_sort_paths.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2
    for n in list(range(0, 40)) + [100, 257, 1000, 5000]:
        for levels in (1, 2, 3, 50):
            score = (rng.integers(0, levels, n) * 0.5).astype(np.float32)
            ham = rng.integers(0, max(1, levels // 2 + 1), n).astype(np.int32)
            start = rng.integers(0, max(1, levels), n).astype(np.int32)
            a = np.zeros(n, np.int32)
            b = np.zeros(n, np.int32)
            hl.ht_sort_paths(_p(score), _p(ham), _p(start), n, _p(a), _p(b))
            assert np.array_equal(a, b), (n, levels)
    # adversarial: organ-pipe and sawtooth patterns push introsort towards its heap-sort fallback
    for n in (3000, 20000):
        for pat in (np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]), np.arange(n) % 7, np.zeros(n)):
            score = pat.astype(np.float32)
            z = np.zeros(score.size, np.int32)
            a = np.zeros(score.size, np.int32)
            b = np.zeros(score.size, np.int32)
            hl.ht_sort_paths(_p(score), _p(z), _p(z), score.size, _p(a), _p(b))
            assert np.array_equal(a, b)


@pytest.mark.parametrize("flat", [0, 1])
@pytest.mark.parametrize("force_scratch", [0, 1])
@pytest.mark.parametrize("name", ["multi_pe", "ties_se", "long", "sync_se", "sync_pe", "flags_se", "ragged_pe", "acc_prune_se", "acc_lvl1_se", "ont_ragged", "redund_se"])
def test_score_core_matches_oracle(hl, name, force_scratch, flat, tmp_path):
    from metabuli_b200 import _ffi
    sdb, reads, seq_mode = synth_cases.build(name)
    odb = oracle.OracleDb.from_synth(sdb)
    hl.ht_set_syncmer(odb.smer_len)                 # syncmer databases: paths may skip up to 8 - s codons, votes per 3 (8 - s) nt
    ov, oq, cov1, cov2 = oracle.extract(*reads, kmer_format=2, syncmer=1 if odb.smer_len else 0, smer_len=odb.smer_len or 5)
    sv, sq = oracle.sort_kmers(ov, oq)
    m = oracle.sort_matches(odb.match(sv, sq))
    c2 = cov2 if seq_mode == 2 else None
    fl = synth_cases.oracle_flags(name)
    fl.pop("lineage", None)                             # a column of the TSV, not a scoring flag
    acc = fl["accession_level"]                         # loadDbParameters (common.cpp:101-108)
    if sdb.database.params.accession_level_db == 1 and acc == 0:
        acc = 2
    if sdb.database.params.accession_level_db != 1 and acc == 1:
        acc = 0
    fl["accession_level"] = acc
    ores, opairs = odb.score(m, cov1, c2, seq_mode=seq_mode, **fl)
    t = sdb.database.tax
    t2s = np.ascontiguousarray(sdb.database.taxid2species)
    tx = _ffi.Taxonomy(t.max_nodes, t.max_taxid, t.eukaryota, _p(t.D), _p(t.E), _p(t.L), _p(t.H), _p(t.M), t.M_k, _p(t.node_taxid),
                       _p(t.node_parent), _p(t.node_prune), _p(t.node_rank), _p(t2s))
    n = cov1.size
    res = np.zeros(n, dtype=_ffi.RESULT_DTYPE)
    pairs = np.zeros((max(16, opairs.shape[0] + 16), 2), dtype=np.int32)
    used = C.c_size_t(0)
    hl.ht_score.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                            C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    mm = np.ascontiguousarray(m)
    rc = hl.ht_score(_p(mm), mm.size, n, _p(cov1), _p(c2), C.byref(tx), seq_mode, fl["min_score"], fl["min_sp_score"], fl["tie_ratio"],
                     fl["min_cons"], fl["min_cons_euk"], acc, 2, force_scratch, flat, _p(res), _p(pairs),
                     pairs.shape[0], C.byref(used))
    assert rc == 0
    for f in ("classification", "query_length", "taxcnt_len", "is_classified", "taxcnt_begin"):
        assert np.array_equal(res[f], ores[f]), f
    assert np.array_equal(res["score"].view(np.uint32), ores["score"].view(np.uint32))
    assert np.array_equal(pairs[: used.value], opairs)
    odb.close()
