"""A scalar model of the per-lane logic of the experimental direct merge (k3_merge.cu, kDirect: group walk with the
predecessor-sum shortcut, sums of the first 16 candidates packed in 4 bits, threshold min(2*min, 7), survivor walk) against the
oracle's matcher on a synthetic index — checks the algorithm the kernel transcribes, not the kernel (that needs a GPU:
tests/test_gpu_synth.py::test_experimental_kernels)."""
import ctypes as C

import numpy as np

import oracle
import shard_oracle
import synth_cases


def _ham(a, b):
    return int(oracle.lib().orc_hamming_sum(int(a), int(b)))


def test_direct_lane_logic_equals_the_oracle_matcher():
    sdb, reads, seq_mode = synth_cases.build("ties_se")
    odb = oracle.OracleDb.from_synth(sdb)
    v, q, _, _ = oracle.extract(*reads)
    sv, sq = oracle.sort_kmers(v, q)
    want = odb.match(sv, sq)
    vals, _ = shard_oracle.decode_stream(np.asarray(sdb.database.diff_idx))
    info = np.asarray(sdb.database.info)
    nk = vals.size - 1                                        # Q1: the numerically last k-mer is never a candidate
    aa = (vals[:nk] >> np.uint64(24))
    keep = ((sq >> np.uint64(32)) & np.uint64(0x1FFFFFFF)) != 0
    qv, qi = sv[keep], sq[keep]
    g0s = np.searchsorted(aa, qv >> np.uint64(24), side="left")
    got = []
    sub = np.random.default_rng(1).choice(qv.size, size=60000, replace=False)       # a sample keeps the pure-Python loop short
    for idx in sub:
        value, qinfo, g0 = int(qv[idx]), int(qi[idx]), int(g0s[idx])
        q40, qd = value >> 24, value & 0xFFFFFF
        if g0 >= nk or int(aa[g0]) != q40:
            continue
        end, best, packed, prev_td, prev_sum = g0, 255, 0, -1, 0
        while end < nk and int(aa[end]) == q40:               # pass 1
            td = int(vals[end]) & 0xFFFFFF
            s = prev_sum if td == prev_td else (0 if td == qd else _ham(qd, td))
            prev_td, prev_sum = td, s
            best = min(best, s)
            if end - g0 < 16:
                packed |= min(s, 15) << (4 * (end - g0))
            end += 1
        limit = min(2 * best, 7)
        for j in range(g0, end):                              # pass 2
            td = int(vals[j]) & 0xFFFFFF
            s = (packed >> (4 * (j - g0))) & 15 if j - g0 < 16 else (0 if td == qd else _ham(qd, td))
            if s <= limit:
                got.append((qinfo, int(info[j]), td, s))
    sel = np.isin(want["qinfo"], qi[sub])
    w = sorted((int(a), int(b), int(c), int(d)) for a, b, c, d in zip(want["qinfo"][sel], want["target_id"][sel], want["dna_encoding"][sel], want["hamming"][sel]))
    # a qinfo can occur for several sampled positions only once (qinfo is unique per query k-mer)
    assert sorted(got) == w and len(w) > 5000
    odb.close()
