"""The range-wise index generator behind the 40 GiB benchmarks (synth.build_db_parts): walking the value space one range at a
time must give the very stream a whole-index build gives, and the ranges taken as shards (each a stream of its own, first delta
relative to 0) must tile it — otherwise the 40 GiB numbers would be measured on a different kind of index than the 8 GiB ones."""
import numpy as np

from metabuli_b200 import synth


def _decode(d, base=0):
    ends = np.nonzero(d & 0x8000)[0]
    out, v, start = [], base, 0
    for e in ends:
        acc = 0
        for x in d[start:e + 1]:
            acc = (acc << 15) | int(x & 0x7FFF)
        v += acc
        out.append(v)
        start = e + 1
    return out


def test_range_wise_build_equals_whole_build():
    kw = dict(genera=5, species_per_genus=4, strains_per_species=2, codons=1200, seed=3)
    whole = synth.make_db(**kw)
    for parts in (2, 5):
        p = synth.make_db(parts=parts, **kw)
        assert np.array_equal(p.database.diff_idx, whole.database.diff_idx)
        assert np.array_equal(p.database.info, whole.database.info)
        assert len(p.range_cuts) == parts - 1 and all(c & 0xFFFFFF == 0 for c in p.range_cuts)      # amino-acid-group aligned


def test_ranges_as_shards_tile_the_index():
    kw = dict(genera=5, species_per_genus=4, strains_per_species=2, codons=1200, seed=3)
    whole = synth.make_db(**kw)
    want = _decode(whole.database.diff_idx)
    got, infos = [], []
    for lo, hi in ((0, 2), (2, 3), (3, 6)):
        s = synth.make_db(parts=6, part=(lo, hi), **kw)
        vals = _decode(s.database.diff_idx)
        assert not vals or vals[0] >= s.shard_first_value
        assert s.shard_is_tail == (hi == 6)
        got += vals
        infos.append(s.database.info)
    assert got == want
    assert np.array_equal(np.concatenate(infos), whole.database.info)
