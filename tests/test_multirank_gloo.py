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
        # 4. peer-memory exchange set-up: the ranks must agree.  Stub handles stand in for the library (no GPU
        #    here): when one rank cannot map its peers, every rank falls back to the NCCL exchange together.
        from fos_b200 import _lib as flib

        class StubLib:
            def __init__(self, fail_import):
                self.fail_import = fail_import
                self.imported = None

            def fos_comm_p2p_export(self, h, arr):
                for k in range(64):
                    arr[k] = (rank * 64 + k) % 251
                return 0

            def fos_comm_p2p_import(self, h, buf):
                self.imported = bytes(buf)
                return -5 if self.fail_import else 0

        class StubHandle:
            def __init__(self, fail_import):
                self.L, self.h, self.options = StubLib(fail_import), None, {}

            def ck(self, rc):
                if rc != 0:
                    raise flib.FosError(rc, "stub failure")

            def set_option(self, k, v):
                self.options[k] = v

        h_ok = StubHandle(False)
        all_ok = parallel.enable_p2p_exchange(h_ok, rank, world, dist)
        table_ok = h_ok.L.imported == bytes([(r * 64 + k) % 251 for r in range(world) for k in range(64)])
        h_bad = StubHandle(fail_import=(rank == 1))
        mixed = parallel.enable_p2p_exchange(h_bad, rank, world, dist)
        fell_back = (h_bad.options.get("exchange_impl") == 0) if rank == 0 else (h_bad.options == {})
        q.put((rank, ok_id, err, (b0, cnt), (bb, bc), all_ok and table_ok, (not mixed) and fell_back))
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
    assert all(r[5] for r in res), "peer-memory handle table was not gathered in rank order"
    assert all(r[6] for r in res), "ranks did not fall back to NCCL together"
