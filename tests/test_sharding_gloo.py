"""CPU test of the multi-GPU plumbing with world_size = 2 over gloo: SNP range partition, the single
broadcast of the Step-1 vectors, and the gather of per-rank result blocks.  The per-rank
compute is a deterministic stand-in (the CUDA path needs a GPU); bench.py --gpus N uses the same helpers."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spagxecct_b200.sharded import run_sharded, shard_range


def test_shard_range_partitions_exactly():
    for m in (0, 1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [shard_range(m, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == m
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, m, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1000
    g = torch.Generator().manual_seed(5)
    R = torch.randn(n, generator=g, dtype=torch.float64) if rank == 0 else torch.zeros(n, dtype=torch.float64)
    E = torch.randn(n, generator=g, dtype=torch.float64) if rank == 0 else torch.zeros(n, dtype=torch.float64)

    def compute(lo, hi, vecs):
        r, e = vecs
        j = np.arange(lo, hi, dtype=np.float64)[:, None]
        return j * 10 + np.arange(3)[None, :] + float((r * e).sum())   # depends on the broadcast vectors

    out = run_sharded(dist, rank, world, m, 3, [R, E], compute)
    if rank == 0:
        q.put((out, float((R * E).sum())))
    dist.destroy_process_group()


def test_two_rank_broadcast_compute_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    m = 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, c = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(m, dtype=np.float64)[:, None] * 10 + np.arange(3)[None, :] + c
    assert np.array_equal(out, want)
