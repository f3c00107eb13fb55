"""mbl_plan_shards (host-only entry point of the C-ABI): the cuts cover the index, sit on amino-acid-group starts, carry the
right decode base, and the oracle matching shard by shard finds exactly the matches it finds on the whole index."""
import numpy as np
import pytest

import oracle
import shard_oracle
import synth_cases

AA = np.uint64(0xFFFFFFFFFF000000)


def _db(split_num, codons=1500, kmer_format=2):
    from metabuli_b200 import synth
    return synth.make_db(genera=4, species_per_genus=3, strains_per_species=2, codons=codons, seed=31, split_num=split_num,
                         kmer_format=kmer_format)


@pytest.mark.parametrize("split_num,n_shards", [(4096, 2), (4096, 8), (64, 3), (2, 4), (4096, 1)])
def test_cuts_are_group_starts_and_cover_the_index(split_num, n_shards):
    from metabuli_b200 import sharded
    sdb = _db(split_num)
    d = sdb.database
    shards = sharded.plan_shards(d, n_shards)
    vals, starts = shard_oracle.decode_stream(np.asarray(d.diff_idx))
    assert vals.size == d.info.size
    assert shards[0].diff_begin == 0 and shards[0].info_begin == 0 and shards[0].first_value == 0 and shards[0].base_value == 0
    non_empty = [s for s in shards if s.info_end > s.info_begin]
    assert non_empty[-1].diff_end == d.diff_idx.size and non_empty[-1].info_end == d.info.size
    assert [s.holds_db_tail for s in non_empty] == [0] * (len(non_empty) - 1) + [1]
    for a, b in zip(non_empty[:-1], non_empty[1:]):
        assert a.diff_end == b.diff_begin and a.info_end == b.info_begin
    for s in non_empty[1:]:
        k = int(s.info_begin)
        assert int(starts[k]) == s.diff_begin
        assert int(vals[k]) == s.first_value and int(vals[k - 1]) == s.base_value
        assert (vals[k] & AA) != (vals[k - 1] & AA)                       # a new amino-acid group starts here
    for s in shards:
        if s.info_end == s.info_begin:
            assert s.first_value == 0xFFFFFFFFFFFFFFFF and not s.holds_db_tail
    if n_shards > 1 and split_num >= 64:
        sizes = [2 * (s.diff_end - s.diff_begin) + 4 * (s.info_end - s.info_begin) for s in shards]
        assert max(sizes) < 1.5 * sum(sizes) / n_shards                  # near-equal bytes


def test_more_shards_than_groups():
    from metabuli_b200 import _ffi, sharded
    import ctypes as C
    # three k-mers in two amino-acid groups
    vals = [(5 << 24) | 1, (5 << 24) | 9, (7 << 24) | 2]
    diff, prev = [], 0
    for v in vals:
        diff += shard_oracle.encode_delta(v - prev)
        prev = v

    class D:
        diff_idx = np.array(diff, dtype=np.uint16); info = np.array([1, 1, 1], dtype=np.int32); split = np.zeros(3, dtype=np.uint64)
    shards = sharded.plan_shards(D, 4)
    assert [int(s.info_end - s.info_begin) for s in shards] == [2, 1, 0, 0]
    assert shards[1].first_value == vals[2] and shards[1].base_value == vals[1] and shards[1].holds_db_tail == 1


@pytest.mark.parametrize("name,n_shards", [("multi_se", 2), ("multi_se", 5), ("format1_pe", 3)])
def test_oracle_matches_shard_by_shard(name, n_shards):
    """Union over the shards of oracle.match(shard DB, the queries routed to it) == oracle.match(whole DB, all queries)."""
    from metabuli_b200 import sharded
    sdb, reads, seq_mode = synth_cases.build(name)
    shards = sharded.plan_shards(sdb.database, n_shards)
    v, q, _, _ = oracle.extract(*reads, kmer_format=sdb.database.params.kmer_format)
    sv, sq = oracle.sort_kmers(v, q)
    full = oracle.OracleDb.from_synth(sdb)
    want = full.match(sv, sq)
    first = np.array([int(s.first_value) for s in shards], dtype=np.uint64) & AA
    keep = ((sq >> np.uint64(32)) & np.uint64(0x1FFFFFFF)) != 0
    sid = np.searchsorted(first, sv & AA, side="right") - 1
    got = []
    for r, s in enumerate(shards):
        odb = shard_oracle.shard_oracle_db(sdb, s)
        sel = keep & (sid == r)
        if odb is None:
            assert not sel.any()
            continue
        got.append(odb.match(sv[sel], sq[sel]))
        odb.close()
    got = np.concatenate(got)
    order = list(want.dtype.names)
    assert np.array_equal(np.sort(got, order=order), np.sort(want, order=order))
    assert want.size > 1000
    full.close()


@pytest.mark.parametrize("db,n_shards", [("in", 2), ("in", 8), ("ex", 3)])
def test_cuts_on_the_reference_built_fixture(db, n_shards, fixtures_dir):
    """The regression fixture databases were written by the reference's own `build` (diffIdx, info and the 4096-entry `split`
    file): the planner must produce valid cuts from those checkpoints too, and the oracle must find the same matches shard by
    shard as on the whole index."""
    import os
    from metabuli_b200 import load_database, read_fastx, sharded
    d = load_database(os.path.join(fixtures_dir, f"db_{db}"))
    shards = sharded.plan_shards(d, n_shards)
    vals, starts = shard_oracle.decode_stream(np.asarray(d.diff_idx))
    assert vals.size == d.info.size
    non_empty = [s for s in shards if s.info_end > s.info_begin]
    assert len(non_empty) == n_shards
    assert non_empty[0].info_begin == 0 and non_empty[-1].info_end == d.info.size and non_empty[-1].diff_end == d.diff_idx.size
    for a, b in zip(non_empty[:-1], non_empty[1:]):
        assert a.diff_end == b.diff_begin and a.info_end == b.info_begin
    for s in non_empty[1:]:
        k = int(s.info_begin)
        assert int(starts[k]) == s.diff_begin and int(vals[k]) == s.first_value and int(vals[k - 1]) == s.base_value
        assert (vals[k] & AA) != (vals[k - 1] & AA)
    sizes = [2 * (s.diff_end - s.diff_begin) + 4 * (s.info_end - s.info_begin) for s in shards]
    assert max(sizes) < 1.3 * sum(sizes) / n_shards
