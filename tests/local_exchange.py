"""TEST INFRASTRUCTURE: an in-process all-to-all between threads, so several ranks of the index-sharded mode can run on ONE
GPU (each rank = a thread with its own mbl context holding one shard).  Same interface as sharded.DistExchange."""
import threading

import torch


class LocalWorld:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.box = [[None] * world for _ in range(world)]      # box[src][dst]

    def exchange(self, rank):
        return LocalExchange(self, rank)


class LocalExchange:
    def __init__(self, w, rank):
        self.w, self.rank, self.world = w, rank, w.world

    def _a2a(self, parts):
        w = self.w
        for dst in range(self.world):
            w.box[self.rank][dst] = parts[dst]
        w.barrier.wait()
        got = [w.box[src][self.rank] for src in range(self.world)]
        w.barrier.wait()
        return got

    same_process = True

    def count_matrix(self, send_counts):
        import numpy as np
        rows = self._a2a([[int(x) for x in send_counts]] * self.world)
        return np.array(rows, dtype=np.int64)

    def all_gather_bytes(self, b):
        return self._a2a([bytes(b)] * self.world)

    def barrier(self):
        self.w.barrier.wait()

    def all_gather_tensor(self, t):
        return self._a2a([t] * self.world)

    def rows(self, send, send_counts, recv_counts):
        parts = list(torch.split(send, [int(x) for x in send_counts], dim=0))
        got = self._a2a([p.clone() for p in parts])
        assert [int(g.shape[0]) for g in got] == [int(x) for x in recv_counts]
        out = torch.cat(got, dim=0) if got else send[:0]
        if out.is_cuda:
            torch.cuda.synchronize(out.device)
        return out
