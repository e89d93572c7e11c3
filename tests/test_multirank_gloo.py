"""world_size-2 `gloo` tests (CPU) of the N>1 host logic: the NCCL-id rendezvous helper and the
row-sharded formulation of the dual mat-vec (each rank's local pass + all-reduce == the full
pass), using the oracle's arithmetic as the per-rank compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import fos_b200  # noqa: F401
    from fos_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. rendezvous of the 128-byte communicator id (rank 0 creates, all receive)
        token = bytes(range(128))
        got = parallel.exchange_comm_id(rank, lambda: token, dist)
        ok_id = got == token
        # 2. row-sharded dual mat-vec: local rows -> partial A'W (all-reduce) + owned rows of A X
        m, n = 100, 37
        rng = np.random.default_rng(0)
        A = rng.standard_normal((m, n))
        X = rng.standard_normal((n, 2))
        W = rng.standard_normal((m, 2))
        b0, cnt = parallel.row_shard(m, rank, world)
        buf = np.zeros((n + m, 2))
        buf[:n] = A[b0:b0 + cnt].T @ W[b0:b0 + cnt]
        buf[n + b0:n + b0 + cnt] = A[b0:b0 + cnt] @ X
        t = torch.from_numpy(buf)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        full = np.vstack([A.T @ W, A @ X])
        err = float(np.abs(t.numpy() - full).max())
        # 3. batch split: disjoint cover
        bb, bc = parallel.batch_shard(11, rank, world)
        q.put((rank, ok_id, err, (b0, cnt), (bb, bc)))
    finally:
        dist.destroy_process_group()


def test_two_rank_rendezvous_and_sharded_matvec():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res)
    assert all(r[2] < 1e-12 for r in res)
    assert res[0][3] == (0, 64) and res[1][3] == (64, 36)
    assert res[0][4] == (0, 6) and res[1][4] == (6, 5)
