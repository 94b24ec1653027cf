"""CPU, world_size 2, gloo: the host-side logic of the N>1 path — contiguous sharding of independent
measurements (identical synthetic data under any sharding), max-over-ranks timing, rank-ordered
gather of per-measurement metrics, and the flat gradient all-reduce of the training step."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deqsci_b200 import distributed as D
    import bench
    n_total = 5
    lo, hi = D.shard_range(n_total, rank, world)
    y, phi, _ = bench.synthetic_batch(lo, hi - lo)
    # checksum of every measurement this rank owns, gathered in rank order
    sums = D.gather_floats([float(y[i].double().sum()) for i in range(hi - lo)])
    slow = D.max_over_ranks(10.0 + rank)
    lin = torch.nn.Linear(3, 2)
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    n = D.allreduce_mean_gradients(lin.parameters())
    ret[rank] = {"range": (lo, hi), "sums": sums, "slow": slow, "n": n,
                 "grads": [float(p.grad.flatten()[0]) for p in lin.parameters()]}
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["range"] == (0, 3) and r1["range"] == (3, 5)          # contiguous, sizes differ by <= 1
    sys.path.insert(0, ROOT)
    import bench
    y, _, _ = bench.synthetic_batch(0, 5)
    want = [float(y[i].double().sum()) for i in range(5)]
    assert r0["sums"] == want and r1["sums"] == want               # same data as an unsharded run
    assert r0["slow"] == r1["slow"] == 11.0                        # max over ranks
    assert r0["n"] == 8 and r0["grads"] == [1.5, 1.5] and r1["grads"] == [1.5, 1.5]


def test_shard_range_covers_everything():
    from deqsci_b200.distributed import shard_range
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
