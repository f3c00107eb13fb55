"""TEST INFRASTRUCTURE: an oracle-backed stand-in for metabuli_b200.sharded.ShardedClassifier (same three phase methods,
CPU tensors) so the exchange logic of the index-sharded mode can run over gloo without a GPU, and helpers that turn one
shard of an index into a stand-alone database the oracle can read."""
import numpy as np
import torch

import oracle

AA_MASK = np.uint64(0xFFFFFFFFFF000000)
SEQ_MASK = np.uint64(0x1FFFFFFF)


def encode_delta(delta: int):
    """IndexCreator::getDiffIdx (IndexCreator.cpp:874-892): 15-bit groups, most significant first, end flag on the last."""
    out = [0x8000 | (delta & 0x7FFF)]
    delta >>= 15
    while delta:
        out.append(delta & 0x7FFF)
        delta >>= 15
    return out[::-1]


def decode_stream(diff, base=0):
    """values of every k-mer of a fragment stream + the fragment index each one starts at (pure Python: small DBs only)."""
    vals, starts = [], []
    v, d, s = base, 0, 0
    for i, f in enumerate(diff.tolist()):
        d = (d << 15) | (f & 0x7FFF)
        if f & 0x8000:
            v += d
            vals.append(v)
            starts.append(s)
            d, s = 0, i + 1
    return np.array(vals, dtype=np.uint64), np.array(starts, dtype=np.int64)


def shard_arrays(database, shard):
    """(diff, info) of a stand-alone index holding exactly the shard's k-mers.  The first delta is re-encoded from 0 and a
    shard that does not end the DB gets a sentinel k-mer in a new amino-acid group, because the oracle (like the reference)
    never matches the numerically last k-mer of the file it reads (Q1)."""
    d = np.asarray(database.diff_idx[int(shard.diff_begin):int(shard.diff_end)])
    info = np.asarray(database.info[int(shard.info_begin):int(shard.info_end)])
    if info.size == 0:
        return np.zeros(0, np.uint16), np.zeros(0, np.int32)
    first_len = int(np.argmax((d & 0x8000) != 0)) + 1
    first_delta = 0
    for f in d[:first_len].tolist():
        first_delta = (first_delta << 15) | (f & 0x7FFF)
    first_value = int(shard.base_value) + first_delta
    head = np.array(encode_delta(first_value), dtype=np.uint16)
    parts, infos = [head, d[first_len:]], [info]
    if not shard.holds_db_tail:
        parts.append(np.array(encode_delta(1 << 25), dtype=np.uint16))
        infos.append(info[:1])
    return np.concatenate(parts), np.concatenate(infos).astype(np.int32)


def shard_oracle_db(sdb, shard):
    import ctypes as C
    d = sdb.database
    diff, info = shard_arrays(d, shard)
    if info.size == 0:
        return None
    split = np.zeros(3, dtype=np.uint64)
    blob = np.frombuffer(sdb.taxonomy_blob, dtype=np.uint8)
    tl = np.ascontiguousarray(sdb.taxid_list, dtype=np.int32)
    err = C.create_string_buffer(512)
    h = oracle.lib().orc_db_from_arrays(oracle._p(diff), diff.size, oracle._p(info), info.size, oracle._p(split), 1, oracle._p(blob), blob.size,
                                        oracle._p(tl), tl.size, d.params.kmer_format, d.params.skip_redundancy, err, 512)
    if not h:
        raise RuntimeError(err.value.decode())
    return oracle.OracleDb(None, handle=h)


class OraclePhases:
    def __init__(self, sdb, shards, rank, seq_mode):
        self.sdb, self.shards, self.rank, self.seq_mode = sdb, shards, rank, seq_mode
        self.first = np.array([int(s.first_value) for s in shards], dtype=np.uint64) & AA_MASK
        self.odb = shard_oracle_db(sdb, shards[rank])
        self.full = oracle.OracleDb.from_synth(sdb)        # scoring only needs the taxonomy

    def phase_extract(self, b1, o1, b2, o2, seq_base):
        v, q, self.cov1, self.cov2 = oracle.extract(b1, o1, b2, o2, kmer_format=self.sdb.database.params.kmer_format)
        self.seq_base = seq_base
        keep = ((q >> np.uint64(32)) & SEQ_MASK) != 0
        v, q = v[keep], q[keep] + (np.uint64(seq_base) << np.uint64(32))
        sid = np.searchsorted(self.first, v & AA_MASK, side="right") - 1
        order = np.argsort(sid, kind="stable")
        counts = np.bincount(sid, minlength=len(self.shards)).tolist()
        self._packed = (torch.from_numpy(v[order].view(np.int64).copy()), torch.from_numpy(q[order].view(np.int64).copy()))
        return counts

    def pack_kmers(self, total):
        assert int(self._packed[0].numel()) == total
        return self._packed

    def phase_match(self, rv, rq, owner_first_read):
        v = rv.numpy().view(np.uint64)
        q = rq.numpy().view(np.uint64)
        n_owners = len(owner_first_read) - 1
        if v.size == 0 or self.odb is None:
            self._packed_m = torch.zeros((0, 3), dtype=torch.int64)
            return [0] * n_owners
        sv, sq = oracle.sort_kmers(v, q)
        m = self.odb.match(sv, sq)
        seq = ((m["qinfo"] >> np.uint64(32)) & SEQ_MASK).astype(np.int64) - 1
        own = np.searchsorted(np.asarray(owner_first_read, dtype=np.int64), seq, side="right") - 1
        order = np.argsort(own, kind="stable")
        counts = np.bincount(own, minlength=n_owners).tolist()
        rows = np.ascontiguousarray(m[order]).view(np.int64).reshape(-1, 3)
        self._packed_m = torch.from_numpy(rows.copy())
        return counts

    def pack_matches(self, total):
        assert int(self._packed_m.shape[0]) == total
        return self._packed_m

    def phase_score(self, rm):
        m = rm.numpy().reshape(-1).view(oracle.MATCH_DTYPE).copy()
        m["qinfo"] -= np.uint64(self.seq_base) << np.uint64(32)
        ms = oracle.sort_matches(m)
        c2 = self.cov2 if self.seq_mode == 2 else None
        return self.full.score(ms, self.cov1, c2, seq_mode=self.seq_mode)
