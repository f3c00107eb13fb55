"""The N>1 host path (replica mode: reads sharded across ranks, results gathered on rank 0) exercised with
world_size 2 on CPU over gloo.  The per-rank classify is stubbed with the oracle — here the thing under test is
the sharding / gather logic of metabuli_b200/multigpu.py, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    import oracle
    import synth_cases
    from metabuli_b200 import multigpu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sdb, reads, seq_mode = synth_cases.build("multi_pe")
    odb = oracle.OracleDb.from_synth(sdb)

    def classify(b1, o1, b2, o2):
        _, res, _, _ = odb.classify_arrays(b1, o1, b2, o2, seq_mode=seq_mode, threads=1)
        pairs = np.zeros((int(res["taxcnt_len"].sum()), 2), np.int32)
        res["taxcnt_begin"] = np.concatenate([[0], np.cumsum(res["taxcnt_len"])[:-1]]).astype(np.uint32)
        return res, pairs

    out = multigpu.classify_sharded(classify, *reads, rank=rank, world=world, dist=dist)
    if rank == 0:
        full, _ = classify(*reads)
        res, pairs = out
        ok = (np.array_equal(res["classification"], full["classification"]) and np.array_equal(res["score"], full["score"])
              and np.array_equal(res["taxcnt_begin"], full["taxcnt_begin"]) and pairs.shape[0] == int(full["taxcnt_len"].sum()))
        q.put(bool(ok))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from metabuli_b200.multigpu import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
