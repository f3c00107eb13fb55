"""Index-sharded multi-GPU mode: diffIdx/info range-partitioned over the ranks at
amino-acid-group boundaries, query metamers sent to the owning shard and matches sent back to the read owner with two
all-to-all exchanges over NCCL (torch.distributed is the plumbing; every compute phase runs in libmetabuli_b200.so).

The reference has no counterpart: its OpenMP threads share one diffIdx and start at one of the `split` checkpoints each
(KmerMatcher.cpp:156-217, Kmer.h:111-119).  The same checkpoints are where the index is cut here (mbl_plan_shards).

    rank r:  phase_extract(own reads)  --a2a #1: (value, qinfo)-->  phase_match(own shard)  --a2a #2: 24-B rows-->  phase_score

`classify_index_sharded` drives one batch through the phases; `phases` is any object with the three phase methods
(ShardedClassifier = the CUDA path; the CPU tests plug in an oracle-backed stand-in to exercise the exchange logic over gloo).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


# ---- shard planning (host only) ------------------------------------------------------------------------------------
def plan_shards(db, n_shards: int):
    """mbl_plan_shards over a dbio.Database -> list of _ffi.Shard (works without a GPU)."""
    lib = _ffi.load_library()
    diff = np.ascontiguousarray(db.diff_idx)
    info = np.ascontiguousarray(db.info)
    split = np.ascontiguousarray(db.split)
    dbs = _ffi.Db(diff.ctypes.data_as(C.c_void_p), diff.size, info.ctypes.data_as(C.c_void_p), info.size,
                  split.ctypes.data_as(C.c_void_p), split.size // 3)
    out = (_ffi.Shard * n_shards)()
    rc = lib.mbl_plan_shards(C.byref(dbs), n_shards, out)
    if rc != _ffi.MBL_OK:
        raise _ffi.MblError(rc, "mbl_plan_shards failed")
    return [out[i] for i in range(n_shards)]


def shard_first_values(shards) -> np.ndarray:
    return np.array([int(s.first_value) for s in shards], dtype=np.uint64)


# ---- exchanges ------------------------------------------------------------------------------------------------------
class DistExchange:
    """Variable-count all-to-all over a torch.distributed process group (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, dist, device):
        self.dist = dist
        self.device = device
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()

    def counts(self, send_counts):
        import torch
        s = torch.tensor([int(x) for x in send_counts], dtype=torch.int64, device=self.device)
        r = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        self.dist.all_to_all_single(r, s)
        return [int(x) for x in r.tolist()]

    def all_gather_int(self, v: int):
        return self.counts([int(v)] * self.world)       # every rank sends its number to every rank

    def rows(self, send, send_counts, recv_counts):
        """send: tensor whose dim 0 is split by send_counts -> tensor of sum(recv_counts) rows."""
        import torch
        recv = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in recv_counts],
                                    input_split_sizes=[int(x) for x in send_counts])
        if send.is_cuda:
            torch.cuda.current_stream(send.device).synchronize()     # the library runs on its own stream
        return recv


# ---- the CUDA phases ------------------------------------------------------------------------------------------------
class _DevArray:
    """Zero-copy view of library-owned device memory for torch (valid until the context's next call)."""

    def __init__(self, ptr: int, shape, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class ShardedClassifier:
    """The classify path of one rank in index-sharded mode: a Classifier context that holds one shard of the index."""

    def __init__(self, database, opt, shards, rank: int):
        from .classifier import Classifier
        self.rank = rank
        self.shards = shards
        self.first_values = shard_first_values(shards)
        self.clf = Classifier(None, opt, database=database, shard=shards[rank])
        self.lib = self.clf.lib
        self.ctx = self.clf.ctx
        self.device = f"cuda:{opt.device}"
        self._keep = None

    def close(self):
        self.clf.close()

    def _tensor(self, ptr, shape):
        import torch
        n = int(np.prod(shape))
        if n == 0 or not ptr:
            return torch.empty(tuple(shape), dtype=torch.int64, device=self.device)
        return torch.as_tensor(_DevArray(ptr, shape), device=self.device)

    def phase_extract(self, bases1, off1, bases2, off2, seq_base: int):
        batch, self._keep = self.clf.make_batch(bases1, off1, bases2, off2)
        self.clf._n_resident = int(batch.n_reads)
        n = len(self.shards)
        counts = np.zeros(n, dtype=np.uint64)
        pv, pq = C.c_void_p(), C.c_void_p()
        self.clf._check(self.lib.mbl_shard_extract(self.ctx, C.byref(batch), int(seq_base), n, self.first_values.ctypes.data_as(C.c_void_p),
                                                   counts.ctypes.data_as(C.c_void_p), C.byref(pv), C.byref(pq)))
        total = int(counts.sum())
        return self._tensor(pv.value, (total,)), self._tensor(pq.value, (total,)), [int(x) for x in counts]

    def phase_match(self, recv_value, recv_qinfo, owner_first_read):
        own = np.ascontiguousarray(owner_first_read, dtype=np.uint64)
        n_owners = own.size - 1
        counts = np.zeros(n_owners, dtype=np.uint64)
        pm = C.c_void_p()
        n = int(recv_value.numel())
        self.clf._check(self.lib.mbl_shard_match(self.ctx, recv_value.data_ptr() if n else None, recv_qinfo.data_ptr() if n else None, n, n_owners,
                                                 own.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), C.byref(pm)))
        total = int(counts.sum())
        return self._tensor(pm.value, (total, 3)), [int(x) for x in counts]

    def phase_score(self, recv_match):
        n = int(recv_match.shape[0])
        self.clf._check(self.lib.mbl_shard_score(self.ctx, recv_match.data_ptr() if n else None, n))
        return self.clf.download_results()


class SelfExchange:
    """world_size 1: the exchanges are the identity (single-GPU runs of the sharded entry points)."""
    world, rank = 1, 0

    def counts(self, send_counts):
        return [int(x) for x in send_counts]

    def all_gather_int(self, v: int):
        return [int(v)]

    def rows(self, send, send_counts, recv_counts):
        return send


# ---- one batch through the sharded path ------------------------------------------------------------------------------
def classify_index_sharded(phases, exchange, bases1, off1, bases2=None, off2=None, timings: dict | None = None):
    """Every rank calls this with ITS reads; returns (results, taxcnt_pairs) for those reads, bit-identical to a
    single-GPU classify of the same reads against the whole index.  timings (optional) receives wall-clock seconds of the
    phases and exchanges (every phase and exchange ends synchronised) and the bytes this rank put on the wire."""
    import time
    t = [time.perf_counter()]

    def lap(key):
        t.append(time.perf_counter())
        if timings is not None:
            timings[key] = timings.get(key, 0.0) + (t[-1] - t[-2])

    n_reads = int(off1.size - 1)
    per_rank = exchange.all_gather_int(n_reads)
    owner_first_read = np.concatenate([[0], np.cumsum(per_rank)]).astype(np.uint64)
    seq_base = int(owner_first_read[exchange.rank])
    # phase 1 + all-to-all #1: metamers to the shard that owns their amino-acid part
    sv, sq, kc = phases.phase_extract(bases1, off1, bases2, off2, seq_base)
    lap("s_extract")
    rc = exchange.counts(kc)
    rv = exchange.rows(sv, kc, rc)
    rq = exchange.rows(sq, kc, rc)
    lap("s_a2a_kmers")
    # phase 2 + all-to-all #2: matches back to the rank that owns the read
    sm, mc = phases.phase_match(rv, rq, owner_first_read)
    del rv, rq, sv, sq
    lap("s_match")
    rmc = exchange.counts(mc)
    rm = exchange.rows(sm, mc, rmc)
    lap("s_a2a_matches")
    # phase 3
    out = phases.phase_score(rm)
    lap("s_score")
    if timings is not None:
        off_rank = lambda c: int(sum(c)) - int(c[exchange.rank])
        timings["a2a_kmer_bytes"] = timings.get("a2a_kmer_bytes", 0) + 16 * off_rank(kc)
        timings["a2a_match_bytes"] = timings.get("a2a_match_bytes", 0) + 24 * off_rank(mc)
    return out
