"""Run under torchrun (one rank per GPU): row-sharded A + the exchange inside the library (NCCL all-reduce, and the
fused peer-memory kernels) must reproduce the single-GPU iterates.  Exit code 0 = parity holds on
every rank.

  --same-device   every rank uses cuda:0 (rendezvous over gloo, no NCCL communicator: NCCL refuses two ranks on one
                  GPU): the peer-memory exchange between PROCESSES is exercised on a box with a single GPU -- the
                  kernels of the two ranks are time-sliced, each exchange completes when the peer gets its slice.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import fos_b200 as fos  # noqa: E402
from fos_b200 import parallel, problems  # noqa: E402
from fos_b200.model import _cone_arrays, _d, _i32p, _i64p  # noqa: E402
from helpers import rel_err  # noqa: E402


def load_dense(H, P, A_rows, r0, cnt):
    t1, l1 = _cone_arrays(P.constr_cones, P.m, "constraint")
    t2, l2 = _cone_arrays(P.var_cones, P.n, "variable")
    b = np.ascontiguousarray(P.b)
    c = np.ascontiguousarray(P.c)
    A_rows = np.ascontiguousarray(A_rows)
    H.ck(H.L.fos_load_conic_dense(H.h, P.m, P.n, A_rows.ctypes.data_as(C.c_void_p), P.n, 0, r0, cnt, _d(b), _d(c),
                                  len(t1), _i32p(t1), _i64p(l1), len(t2), _i32p(t2), _i64p(l2)))


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    one_dev = "--same-device" in sys.argv
    if one_dev:
        local = 0
    torch.cuda.set_device(local)
    if one_dev:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [(m, n, scale, tol, alg, ex) for ex in (("p2p",) if one_dev else ("nccl", "p2p"))
             for (m, n, scale, tol, alg) in ((1000, 2500, 0.1, 1e-10, "DR"), (2000, 1300, 1.0, 1e-4, "GAPA"),
                                             (333, 4100, 0.1, 1e-10, "FISTA"))]
    for (m, n, scale, tol, alg, ex) in cases:
        P = problems.lasso_like(m, n, seed=7, scale=scale)
        A = np.asarray(P.A)
        r0, cnt = parallel.row_shard(m, rank, world)
        Hs = fos.Handle(local)
        cid = None if one_dev else parallel.exchange_comm_id(rank, parallel.nccl_unique_id, dist)
        parallel.init_comm(Hs, rank, world, cid)
        load_dense(Hs, P, A[r0:r0 + cnt], r0, cnt)
        if ex == "p2p":  # fused peer-memory exchange kernel instead of fold + ncclAllReduce
            assert parallel.enable_p2p_exchange(Hs, rank, world, dist) or not one_dev, "CUDA IPC unavailable"
        H1 = fos.Handle(local)  # unsharded reference on the same device
        load_dense(H1, P, A, 0, m)
        for H in (Hs, H1):
            H.set_algorithm({"DR": fos.DR(0.5), "GAPA": fos.GAPA(), "FISTA": fos.FISTA()}[alg])
            H.set_initial_iterate()
            H.ck(H.L.fos_begin_solve(H.h))
        v = np.random.default_rng(1).standard_normal(2 * (m + n + 1))
        e_kkt = rel_err(Hs.kkt_mul(v), H1.kkt_mul(v))
        worst, flips = 0.0, 0
        for i in range(1, 21):
            # lock-step against the single-GPU handle
            Hs.set_state("x", H1.get_state("x"))
            if H1.info("s1_calls") > 1:
                Hs.set_state("xinit", H1.get_state("xinit"))
            Hs.set_info("s1_calls", H1.info("s1_calls"))
            if alg == "GAPA":
                Hs.set_info("alpha12", H1.info("alpha12"))
            if alg == "FISTA":
                Hs.set_state("fista_y", H1.get_state("fista_y"))
                Hs.set_info("fista_t", H1.info("fista_t"))
            H1.run(i, 1, 5, 1e-9)
            _, _, rec, _ = Hs.run(i, 1, 5, 1e-9)
            flips += Hs.info("cgiter") != H1.info("cgiter")
            worst = max(worst, rel_err(Hs.get_iterate(), H1.get_iterate()))
        good = e_kkt < 1e-12 and worst < tol and flips <= 2
        ok = ok and good
        print(f"rank {rank}/{world} {m}x{n} {alg} [{ex}]: kkt_mul err {e_kkt:.2e}, worst lock-step deviation {worst:.2e}, "
              f"CG count flips {flips} -> {'ok' if good else 'FAIL'}", flush=True)
        # free-running: every rank must take identical decisions (replicated scalars from identical sums)
        Hs.set_initial_iterate()
        Hs.ck(Hs.L.fos_begin_solve(Hs.h))
        done, st, rec, _ = Hs.run(1, 60, 10, 1e-7)
        sig = torch.tensor([float(done), float(st), float(Hs.info("total_cg")), float(np.sum(Hs.get_iterate()))],
                           dtype=torch.float64, device="cpu" if one_dev else "cuda")
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        same = all(torch.equal(sigs[0], t) for t in sigs)
        if not same:
            ok = False
            print(f"rank {rank}: free-running state differs across ranks [{ex}] {sigs}", flush=True)
        del Hs, H1
    flag = torch.tensor([0 if ok else 1], device="cpu" if one_dev else "cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
