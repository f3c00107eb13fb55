"""Replica-mode multi-GPU driver: one process per GPU, the index replicated, the reads of a batch sharded by
contiguous ranges, results gathered on rank 0.  The path has no data-path collective in this
mode; torch.distributed is used for the rendezvous and the gather of the small per-read result records only.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_reads: int, rank: int, world: int):
    """Contiguous [begin, end) of rank's reads; sizes differ by at most one."""
    q, r = divmod(n_reads, world)
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def slice_batch(bases, offsets, begin: int, end: int):
    """Sub-batch [begin, end) of an SoA batch, offsets rebased to zero."""
    o = np.ascontiguousarray(offsets[begin:end + 1])
    b = np.ascontiguousarray(bases[int(o[0]):int(o[-1])])
    return b, (o - o[0]).astype(np.uint64)


def classify_sharded(classify_fn, bases1, off1, bases2=None, off2=None, rank: int = 0, world: int = 1, dist=None):
    """Every rank classifies its shard with classify_fn(b1, o1, b2, o2) -> (results, pairs); rank 0 returns the
    concatenation in read order (taxcnt_begin rebased), other ranks return None."""
    n = off1.size - 1
    b, e = shard_range(n, rank, world)
    s1 = slice_batch(bases1, off1, b, e)
    s2 = slice_batch(bases2, off2, b, e) if bases2 is not None else (None, None)
    res, pairs = classify_fn(s1[0], s1[1], s2[0], s2[1])
    if world == 1 or dist is None:
        return res, pairs
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((res, pairs), gathered, dst=0)
    if rank != 0:
        return None
    out_res, out_pairs, base = [], [], 0
    for r, p in gathered:
        r = r.copy()
        r["taxcnt_begin"] += np.uint32(base)
        base += p.shape[0]
        out_res.append(r)
        out_pairs.append(p)
    return np.concatenate(out_res), np.concatenate(out_pairs) if out_pairs else np.zeros((0, 2), np.int32)
