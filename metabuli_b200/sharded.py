"""Index-sharded multi-GPU mode: diffIdx/info range-partitioned over the ranks at
amino-acid-group boundaries, query metamers sent to the owning shard and matches sent back to the read owner with two
all-to-all exchanges over NCCL (torch.distributed is the plumbing; every compute phase runs in libmetabuli_b200.so).

The reference has no counterpart: its OpenMP threads share one diffIdx and start at one of the `split` checkpoints each
(KmerMatcher.cpp:156-217, Kmer.h:111-119).  The same checkpoints are where the index is cut here (mbl_plan_shards).

    rank r:  phase_extract(own reads)  --a2a #1: (value, qinfo)-->  phase_match(own shard)  --a2a #2: 24-B rows-->  phase_score

Two transports for the exchanges: "peer" — the library's bucket-gather kernels store every row straight into the receiving
rank's buffer over NVLink (CUDA-IPC-mapped peer memory; gather and all-to-all are ONE kernel, then a barrier) — and
"collective" — pack into send buffers + all_to_all_single (NCCL; gloo in the CPU tests).

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
    same_process = False

    def __init__(self, dist, device):
        self.dist = dist
        self.device = device
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()

    def count_matrix(self, send_counts) -> np.ndarray:
        """M[src][dst] = rows rank src sends to rank dst (every rank learns the whole matrix)."""
        import torch
        s = torch.tensor([int(x) for x in send_counts], dtype=torch.int64, device=self.device)
        parts = [torch.zeros(self.world, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        self.dist.all_gather(parts, s)
        return np.array([[int(x) for x in p.tolist()] for p in parts], dtype=np.int64)

    def all_gather_bytes(self, b: bytes):
        out = [None] * self.world
        self.dist.all_gather_object(out, bytes(b))
        return out

    def barrier(self):
        self.dist.barrier()

    def all_gather_tensor(self, t):
        import torch
        parts = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        return parts

    def broadcast_tensor(self, t, src: int):
        import torch
        self.dist.broadcast(t, src=src)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()

    def rows(self, send, send_counts, recv_counts):
        """send: tensor whose dim 0 is split by send_counts -> tensor of sum(recv_counts) rows."""
        import torch
        recv = torch.empty((int(sum(recv_counts)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        self.dist.all_to_all_single(recv, send, output_split_sizes=[int(x) for x in recv_counts],
                                    input_split_sizes=[int(x) for x in send_counts])
        if send.is_cuda:
            torch.cuda.current_stream(send.device).synchronize()     # the library runs on its own stream
        return recv


class SelfExchange:
    """world_size 1: the exchanges are the identity (single-GPU runs of the sharded entry points)."""
    world, rank, same_process = 1, 0, True

    def count_matrix(self, send_counts):
        return np.array([[int(x) for x in send_counts]], dtype=np.int64)

    def all_gather_bytes(self, b):
        return [bytes(b)]

    def barrier(self):
        pass

    def all_gather_tensor(self, t):
        return [t]

    def rows(self, send, send_counts, recv_counts):
        return send


# ---- the CUDA phases ------------------------------------------------------------------------------------------------
class _DevArray:
    """Zero-copy view of library-owned device memory for torch (valid until the context's next phase)."""

    def __init__(self, ptr: int, shape, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class ShardedClassifier:
    """The classify path of one rank in index-sharded mode: a Classifier context that holds one shard of the index."""

    def __init__(self, database, opt, shards, rank: int, total_kmers: int | None = None):
        """database: the whole index (shards[rank] is cut out of it), or — with total_kmers — only this rank's part as a stream of
        its own (shards[rank] then spans it from 0; the other entries of `shards` only carry their first_value)."""
        from .classifier import Classifier
        self.rank = rank
        self.shards = shards
        self.first_values = shard_first_values(shards)
        self.clf = Classifier(None, opt, database=database, shard=shards[rank], total_kmers=total_kmers)
        self.lib = self.clf.lib
        self.ctx = self.clf.ctx
        self.device = f"cuda:{opt.device}"
        self._keep = None
        self.peer_rows = (0, 0)            # capacity (k-mer rows, match rows) of every rank's receive buffers
        self._recv = (0, 0)                # this rank's receive buffers (device pointers)

    def close(self):
        self.clf.close()

    def _tensor(self, ptr, shape):
        import torch
        n = int(np.prod(shape))
        if n == 0 or not ptr:
            return torch.empty(tuple(shape), dtype=torch.int64, device=self.device)
        return torch.as_tensor(_DevArray(ptr, shape), device=self.device)

    def merge_filters(self, exchange):
        """Collective, once after loading: OR the ranks' parts of the amino-acid presence filter together so that phase 1 can
        drop the metamers no shard can match before they are bucketed and sent (4-5x less exchange traffic)."""
        p, nb = C.c_void_p(), C.c_uint64(0)
        self.clf._check(self.lib.mbl_shard_filter(self.ctx, C.byref(p), C.byref(nb)))
        if not p.value or nb.value == 0:
            return False
        mine = self._tensor(p.value, (nb.value // 8,))
        if hasattr(exchange, "broadcast_tensor"):
            # one scratch buffer instead of world copies (the filter of a 40 GiB index is ~11 GB): rank r broadcasts what it
            # has at its turn — possibly already OR-ed with earlier ranks' parts, which changes nothing (OR is idempotent)
            import torch
            tmp = torch.empty_like(mine)
            for r in range(exchange.world):
                if r == exchange.rank:
                    exchange.broadcast_tensor(mine, r)
                else:
                    exchange.broadcast_tensor(tmp, r)
                    self.clf._check(self.lib.mbl_shard_filter_or(self.ctx, tmp.data_ptr(), nb.value, 0))
            del tmp
            torch.cuda.empty_cache()
        else:
            for r, part in enumerate(exchange.all_gather_tensor(mine)):
                if r != exchange.rank:
                    self.clf._check(self.lib.mbl_shard_filter_or(self.ctx, part.data_ptr(), nb.value, 0))
        self.clf._check(self.lib.mbl_shard_filter_or(self.ctx, None, nb.value, 1))
        exchange.barrier()
        return True

    # -- phases ---------------------------------------------------------------------------------------------------
    def phase_extract(self, bases1, off1, bases2, off2, seq_base: int):
        batch, self._keep = self.clf.make_batch(bases1, off1, bases2, off2)
        self.clf._n_resident = int(batch.n_reads)
        n = len(self.shards)
        counts = np.zeros(n, dtype=np.uint64)
        self.clf._check(self.lib.mbl_shard_extract(self.ctx, C.byref(batch), int(seq_base), n, self.first_values.ctypes.data_as(C.c_void_p),
                                                   counts.ctypes.data_as(C.c_void_p)))
        return [int(x) for x in counts]

    def phase_match(self, recv_value, recv_qinfo, owner_first_read):
        own = np.ascontiguousarray(owner_first_read, dtype=np.uint64)
        n_owners = own.size - 1
        counts = np.zeros(n_owners, dtype=np.uint64)
        n = int(recv_value.numel())
        self.clf._check(self.lib.mbl_shard_match(self.ctx, recv_value.data_ptr() if n else None, recv_qinfo.data_ptr() if n else None, n, n_owners,
                                                 own.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p)))
        return [int(x) for x in counts]

    def phase_score(self, recv_match):
        n = int(recv_match.shape[0])
        self.clf._check(self.lib.mbl_shard_score(self.ctx, recv_match.data_ptr() if n else None, n))
        return self.clf.download_results()

    # -- transport A: contiguous send buffers for a collective ----------------------------------------------------
    def pack_kmers(self, total: int):
        pv, pq = C.c_void_p(), C.c_void_p()
        self.clf._check(self.lib.mbl_shard_pack_kmers(self.ctx, C.byref(pv), C.byref(pq)))
        return self._tensor(pv.value, (total,)), self._tensor(pq.value, (total,))

    def pack_matches(self, total: int):
        pm = C.c_void_p()
        self.clf._check(self.lib.mbl_shard_pack_matches(self.ctx, C.byref(pm)))
        return self._tensor(pm.value, (total, 3))

    # -- transport B: peers store into this rank's receive buffers ----------------------------------------------------
    supports_push = True

    def setup_peer_buffers(self, exchange, kmer_rows: int, match_rows: int):
        """Collective: every rank (re)allocates its receive buffers with the same capacity and maps everybody else's."""
        lib = self.lib
        exchange.barrier()
        lib.mbl_shard_detach_peers(self.ctx)
        exchange.barrier()
        dk, dm = C.c_void_p(), C.c_void_p()
        hk = (C.c_uint8 * _ffi.IPC_HANDLE_BYTES)()
        hm = (C.c_uint8 * _ffi.IPC_HANDLE_BYTES)()
        self.clf._check(lib.mbl_shard_recv_buffers(self.ctx, int(kmer_rows), int(match_rows), C.byref(dk), C.byref(dm), hk, hm))
        self._recv = (int(dk.value), int(dm.value))
        mine = bytes(hk) + bytes(hm) + int(dk.value).to_bytes(8, "little") + int(dm.value).to_bytes(8, "little")
        everyone = exchange.all_gather_bytes(mine)
        H = _ffi.IPC_HANDLE_BYTES
        for peer, blob in enumerate(everyone):
            raw_k = int.from_bytes(blob[2 * H:2 * H + 8], "little")
            raw_m = int.from_bytes(blob[2 * H + 8:2 * H + 16], "little")
            if peer == exchange.rank or exchange.same_process:
                rc = lib.mbl_shard_attach_peer(self.ctx, peer, None, None, raw_k, raw_m)
            else:
                bk = (C.c_uint8 * H).from_buffer_copy(blob[:H])
                bm = (C.c_uint8 * H).from_buffer_copy(blob[H:2 * H])
                rc = lib.mbl_shard_attach_peer(self.ctx, peer, bk, bm, None, None)
            self.clf._check(rc)
        self.peer_rows = (int(kmer_rows), int(match_rows))
        exchange.barrier()

    def push_kmers(self, row_off, totals):
        ro = np.ascontiguousarray(row_off, dtype=np.uint64)
        tt = np.ascontiguousarray(totals, dtype=np.uint64)
        self.clf._check(self.lib.mbl_shard_push_kmers(self.ctx, ro.ctypes.data_as(C.c_void_p), tt.ctypes.data_as(C.c_void_p)))

    def push_matches(self, row_off):
        ro = np.ascontiguousarray(row_off, dtype=np.uint64)
        self.clf._check(self.lib.mbl_shard_push_matches(self.ctx, ro.ctypes.data_as(C.c_void_p)))

    def received_kmers(self, total: int):
        return self._tensor(self._recv[0], (total,)), self._tensor(self._recv[0] + 8 * total if total else 0, (total,))

    def received_matches(self, total: int):
        return self._tensor(self._recv[1], (total, 3))


# ---- one batch through the sharded path ------------------------------------------------------------------------------
def classify_index_sharded(phases, exchange, bases1, off1, bases2=None, off2=None, timings: dict | None = None, transport: str = "auto"):
    """Every rank calls this with ITS reads; returns (results, taxcnt_pairs) for those reads, bit-identical to a
    single-GPU classify of the same reads against the whole index.

    transport: "peer" = the gather kernels store into the receivers' buffers over NVLink (fused gather + all-to-all),
    "collective" = pack + all_to_all_single (NCCL / gloo), "auto" = peer when the phases support it.
    timings (optional) receives wall-clock seconds of the phases and exchanges (each ends synchronised) and the bytes this rank
    put on the wire."""
    import time
    t = [time.perf_counter()]

    def lap(key):
        t.append(time.perf_counter())
        if timings is not None:
            timings[key] = timings.get(key, 0.0) + (t[-1] - t[-2])

    rank, world = exchange.rank, exchange.world
    use_peer = transport == "peer" or (transport == "auto" and getattr(phases, "supports_push", False))
    n_reads = int(off1.size - 1)
    per_rank = exchange.count_matrix([n_reads] * world)[:, rank]
    owner_first_read = np.concatenate([[0], np.cumsum(per_rank)]).astype(np.uint64)
    seq_base = int(owner_first_read[rank])

    # phase 1 + all-to-all #1: metamers to the shard that owns their amino-acid part
    kc = phases.phase_extract(bases1, off1, bases2, off2, seq_base)
    K = exchange.count_matrix(kc)                       # K[src][dst]
    lap("s_extract")
    k_tot = K.sum(axis=0)
    if use_peer and int(k_tot.max()) > phases.peer_rows[0]:
        # receive buffers sized from the traffic seen (same decision on every rank: it depends on the shared matrix only)
        phases.setup_peer_buffers(exchange, int(1.25 * k_tot.max()) + 4096, max(phases.peer_rows[1], int(0.6 * k_tot.max()) + 4096))
    if use_peer:
        phases.push_kmers(K[:rank].sum(axis=0), k_tot)
        exchange.barrier()
        rv, rq = phases.received_kmers(int(k_tot[rank]))
    else:
        sv, sq = phases.pack_kmers(int(sum(kc)))
        rv = exchange.rows(sv, kc, K[:, rank])
        rq = exchange.rows(sq, kc, K[:, rank])
        del sv, sq
    lap("s_a2a_kmers")

    # phase 2 + all-to-all #2: matches back to the rank that owns the read
    mc = phases.phase_match(rv, rq, owner_first_read)
    del rv, rq
    Mx = exchange.count_matrix(mc)
    lap("s_match")
    m_tot = Mx.sum(axis=0)
    if use_peer and int(m_tot.max()) > phases.peer_rows[1]:
        phases.setup_peer_buffers(exchange, phases.peer_rows[0], int(1.25 * m_tot.max()) + 4096)
    if use_peer:
        phases.push_matches(Mx[:rank].sum(axis=0))
        exchange.barrier()
        rm = phases.received_matches(int(m_tot[rank]))
    else:
        sm = phases.pack_matches(int(sum(mc)))
        rm = exchange.rows(sm, mc, Mx[:, rank])
        del sm
    lap("s_a2a_matches")

    # phase 3
    out = phases.phase_score(rm)
    lap("s_score")
    if timings is not None:
        timings["a2a_kmer_bytes"] = timings.get("a2a_kmer_bytes", 0) + 16 * (int(sum(kc)) - int(kc[rank]))
        timings["a2a_match_bytes"] = timings.get("a2a_match_bytes", 0) + 24 * (int(sum(mc)) - int(mc[rank]))
    return out
