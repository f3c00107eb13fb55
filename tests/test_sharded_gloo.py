"""The index-sharded N>1 path (metabuli_b200/sharded.py: plan_shards + classify_index_sharded with its two variable-count
all-to-all exchanges) on CPU: world_size 2 over gloo and world_size 3 in-process.  The per-rank compute phases are stubbed
with the oracle (tests/shard_oracle.py) — what is under test is the shard planning, routing, exchange and seqID bookkeeping,
which are the same code the GPUs run.  Results must equal the oracle on the whole index."""
import os
import socket
import sys
import threading

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _expected(sdb, reads, seq_mode):
    import oracle
    odb = oracle.OracleDb.from_synth(sdb)
    v, q, cov1, cov2 = oracle.extract(*reads, kmer_format=sdb.database.params.kmer_format)
    sv, sq = oracle.sort_kmers(v, q)
    m = oracle.sort_matches(odb.match(sv, sq))
    return odb.score(m, cov1, cov2 if seq_mode == 2 else None, seq_mode=seq_mode)


def _same(res, pairs, want_res, want_pairs, lo, hi):
    w = want_res[lo:hi]
    for f in ("classification", "query_length", "taxcnt_len", "is_classified"):
        if not np.array_equal(res[f], w[f]):
            return False
    if not np.array_equal(res["score"].view(np.uint32), w["score"].view(np.uint32)):
        return False
    for i in range(hi - lo):
        a = pairs[int(res["taxcnt_begin"][i]):int(res["taxcnt_begin"][i]) + int(res["taxcnt_len"][i])]
        b = want_pairs[int(w["taxcnt_begin"][i]):int(w["taxcnt_begin"][i]) + int(w["taxcnt_len"][i])]
        if not np.array_equal(a, b):
            return False
    return True


def _worker(rank, world, port, q):
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import shard_oracle
    import synth_cases
    from metabuli_b200 import multigpu, sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sdb, reads, seq_mode = synth_cases.build("multi_pe")
    shards = sharded.plan_shards(sdb.database, world)
    phases = shard_oracle.OraclePhases(sdb, shards, rank, seq_mode)
    n = reads[1].size - 1
    lo, hi = multigpu.shard_range(n, rank, world)
    b1, o1 = multigpu.slice_batch(reads[0], reads[1], lo, hi)
    b2, o2 = multigpu.slice_batch(reads[2], reads[3], lo, hi)
    res, pairs = sharded.classify_index_sharded(phases, sharded.DistExchange(dist, "cpu"), b1, o1, b2, o2)
    want_res, want_pairs = _expected(sdb, reads, seq_mode)
    q.put((rank, bool(_same(res, pairs, want_res, want_pairs, lo, hi)), int(res["is_classified"].sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world_size_2_gloo_index_sharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert [g[1] for g in got] == [True, True]
    assert sum(g[2] for g in got) > 1000


@pytest.mark.parametrize("name,world,idle", [("multi_se", 3, None), ("ragged_se", 4, None), ("multi_se", 3, 1)])
def test_in_process_ranks(name, world, idle):
    import shard_oracle
    import synth_cases
    from local_exchange import LocalWorld
    from metabuli_b200 import multigpu, sharded
    sdb, reads, seq_mode = synth_cases.build(name)
    shards = sharded.plan_shards(sdb.database, world)
    lw = LocalWorld(world)
    want_res, want_pairs = _expected(sdb, reads, seq_mode)
    n = reads[1].size - 1
    ok = [None] * world

    def run(rank):
        phases = shard_oracle.OraclePhases(sdb, shards, rank, seq_mode)
        lo, hi = multigpu.shard_range(n, rank, world)
        if idle is not None:                        # one rank brings no reads but still serves its shard
            cuts = [0] + [n * (r + 1) // world for r in range(world)]
            cuts[idle + 1] = cuts[idle]
            lo, hi = cuts[rank], cuts[rank + 1] if rank + 1 < world else n
        b1, o1 = multigpu.slice_batch(reads[0], reads[1], lo, hi)
        res, pairs = sharded.classify_index_sharded(phases, lw.exchange(rank), b1, o1)
        ok[rank] = res.size == hi - lo and _same(res, pairs, want_res, want_pairs, lo, hi)

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert ok == [True] * world
