"""Measurement of BASELINE.json's configs 3, 4 and 5 on one GPU (config 2 is bench.py's line; the
multi-GPU legs are bench.py --gpus N and scripts/multi_gpu_check.py).  One JSON line per case on stdout.

  python scripts/config_runs.py c5 [--nprob 1024] [--iters 200]
  python scripts/config_runs.py c4 [--d 512] [--iters 30]
  python scripts/config_runs.py c3 [--md 100000] [--nx 20000] [--iters 30]

Timing: CUDA events on the library's stream around the timed call, after a warm-up call; inputs are
resident in HBM.  "cpu" legs time the oracle (single thread, restated reference) on ONE problem /
a bounded sample and are reported next to the GPU number, never used by it.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def timed(torch, stream, fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def run_c5(args):
    """C5: `--nprob` independent NNLS 256 x 512 problems (conic form m = 769, n = 513), FISTA and Dykstra.  Under
    torchrun the batch is split across the ranks with parallel.batch_shard (no collective on the data
    path); problems are generated in chunks of 64 with chunk-indexed seeds, so the batch is the same for every
    rank count.  Timing: barrier, CUDA events on every rank's library stream, max over ranks."""
    import os
    import torch
    import fos_b200 as fos
    from fos_b200 import parallel
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    Btot, rows, cols = args.nprob, 256, 512
    p0, B = parallel.batch_shard(Btot, rank, world)
    m, n = rows + 1 + cols, cols + 1
    CH = 64
    D = torch.empty((B, rows, cols), dtype=torch.float64, device=dev)
    d = torch.empty((B, rows), dtype=torch.float64, device=dev)
    c0 = (p0 // CH) * CH
    while c0 < p0 + B:
        g = torch.Generator(device=dev)
        g.manual_seed(5 * 1000003 + c0 // CH)
        Dc = torch.randn((CH, rows, cols), dtype=torch.float64, device=dev, generator=g) / np.sqrt(cols)
        dc = torch.randn((CH, rows), dtype=torch.float64, device=dev, generator=g)
        lo, hi = max(c0, p0), min(c0 + CH, p0 + B)
        D[lo - p0:hi - p0] = Dc[lo - c0:hi - c0]
        d[lo - p0:hi - p0] = dc[lo - c0:hi - c0]
        c0 += CH
    A = torch.zeros((B, m, n), dtype=torch.float64, device=dev)
    A[:, 0, 0] = -1.0
    A[:, 1:rows + 1, 1:] = -D
    idx = torch.arange(cols, device=dev)
    A[:, rows + 1 + idx, 1 + idx] = -1.0
    del D
    b = np.zeros((B, m))
    b[:, 1:rows + 1] = -d.cpu().numpy()
    c = np.zeros((B, n))
    c[:, 0] = 1.0
    cones1, cones2 = [("SOC", rows + 1), ("NonNeg", cols)], [("Free", n)]
    peak, peak_src = peak_hbm()
    for name, alg in (("FISTA", fos.FISTA()), ("Dykstra", fos.Dykstra())):
        H = fos.Handle(local)
        H.set_option("batch_hybrid", 0 if args.dense_batch else 1)
        if args.batch_ctas:
            H.set_option("batch_ctas", args.batch_ctas)
        H.load_conic_batch((B, m, n), b, c, cones1, cones2, device_ptr=(A.data_ptr(), n, m * n))
        H.set_algorithm(alg)
        stream = torch.cuda.ExternalStream(H.stream(), device=dev)
        H.ck(H.L.fos_begin_solve_batch(H.h))
        W, K = args.warmup, args.iters
        H.run_batch(1, W, 10 ** 9, 1e-5)
        pp0 = H.info_batch("total_passes").sum()
        cg0 = H.info_batch("total_cg").sum()
        if world > 1:
            dist.barrier()
        ms, (done, st, recs) = timed(torch, stream, lambda: H.run_batch(W + 1, K, K, 1e-5))
        passes = H.info_batch("total_passes").sum() - pp0
        cgs = H.info_batch("total_cg").sum() - cg0
        ms_local = ms
        tot = torch.tensor([float(passes), float(cgs), float((done == K).all()), float(np.sum(H.get_iterate_batch()))],
                           dtype=torch.float64, device=dev)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
            dist.all_reduce(tot)
        passes_all, cgs_all, all_done, checksum = [float(x) for x in tot.tolist()]
        bytes_pass = H.info("bytes_per_pass")   # dense rows as FP64 + sparse rows as CSR/CSC entries, per problem
        gbs = passes_all * bytes_pass / (ms / 1e3) / 1e9
        line = {"config": "C5", "n_gpus": world, "algorithm": name, "nprob": Btot, "nprob_per_gpu": B, "m": m, "n": n,
                "batch_ctas": args.batch_ctas, "dense_equivalent_bytes_per_pass": 8.0 * m * n,
                "iterations_timed": K, "ms_total": ms, "ms_rank0": ms_local,
                "problem_iterations_per_s": Btot * K / (ms / 1e3),
                "cg_iterations_per_step": cgs_all / (Btot * K), "passes_over_A_per_step": passes_all / (Btot * K),
                "algorithmic_bytes_per_pass": bytes_pass, "aggregate_gbs": gbs, "per_gpu_gbs": gbs / world,
                "peak_gbs_per_gpu": peak, "frac_per_gpu": gbs / world / peak, "peak_source": peak_src,
                "all_iterated": bool(all_done == world), "iterate_checksum": checksum,
                "check_p_median_rank0": float(np.median([r[-1][1] for r in recs if len(r)])) if K else None}
        if args.cpu and rank == 0:
            from oracle import fos_oracle as fo
            import scipy.sparse as sp
            fo.build()
            A0 = sp.csc_matrix(A[0].cpu().numpy())
            O = fo.OracleConic(c[0], A0, b[0], cones1, cones2)
            O.set_algorithm(name, 1.0, 0.0, 0.0, 0.0, 100)
            O.set_iterate(O.initial_value())
            O.run(1, W, checki=10 ** 9, eps=1e-5)
            t0 = time.perf_counter()
            kk = min(K, 50)
            O.run(W + 1, kk, checki=10 ** 9, eps=1e-5)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": kk / dt, "unit": "problem-iterations/s", "cores": 1, "kind": "port",
                                    "sample": f"oracle (C, CSC) on problem 0, iterations {W + 1}..{W + kk}"}
        if rank == 0:
            print(json.dumps(line), flush=True)
        del H
    if world > 1:
        dist.destroy_process_group()


def run_c4(args):
    import torch
    import fos_b200 as fos
    from fos_b200 import problems
    from helpers import load_conic
    dev = torch.device("cuda", 0)
    d = args.d
    P = problems.sdp_nearest_correlation(d, seed=4)
    H = load_conic(fos, P, storage="sparse")
    H.set_algorithm(fos.GAP(0.8, 1.8, 1.8))
    stream = torch.cuda.ExternalStream(H.stream(), device=dev)
    H.ck(H.L.fos_begin_solve(H.h))
    W, K = args.warmup, args.iters
    H.run(1, W, 100, 1e-5)
    cg0 = H.info("total_cg")
    ms, (done, st, rec, _) = timed(torch, stream, lambda: H.run(W + 1, K, 100, 1e-5))
    cgs = H.info("total_cg") - cg0
    rng = np.random.default_rng(0)
    G = rng.standard_normal((d, d))
    X = problems.svec((G + G.T) / 2)
    _, ms_psd2, sweeps = H.time_psd(np.stack([X, -X]), reps=3)
    line = {"config": "C4", "algorithm": "GAP(0.8,1.8,1.8)", "sdp_order": d, "m": P.m, "n": P.n, "iterations_timed": int(done),
            "ms_per_iteration": ms / max(done, 1), "iterations_per_s": done / (ms / 1e3),
            "cg_iterations_per_step": cgs / max(done, 1), "cold_psd_projection_2x_ms": ms_psd2, "cold_jacobi_sweeps": sweeps,
            "status": int(st)}
    if args.cpu:
        from oracle import fos_oracle as fo
        fo.build()
        O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
        O.set_algorithm("GAP", 0.8, 1.8, 1.8, 0.0, 100)
        O.set_iterate(O.initial_value())
        t0 = time.perf_counter()
        kk = 3
        O.run(1, kk, checki=100, eps=1e-5)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": kk / dt, "unit": "iterations/s", "cores": 1, "kind": "port",
                                "sample": f"oracle (C, Jacobi eigensolver) iterations 1..{kk}"}
    print(json.dumps(line), flush=True)


def run_c3(args):
    """C3: min t s.t. ||D x - d|| <= t, ||x|| <= rho, D md x nx dense; K1 = SOC(md+1) + SOC(nx+1); GAPA().
    Under torchrun the (md+nx+2) x (nx+1) matrix is row-sharded over the ranks."""
    import ctypes as C
    import os
    import torch
    import fos_b200 as fos
    from fos_b200 import parallel
    from fos_b200.model import _cone_arrays, _d, _i32p, _i64p
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    md, nx = args.md, args.nx
    m, n = md + 1 + nx + 1, nx + 1
    r0, cnt = parallel.row_shard(m, rank, world)
    A = torch.zeros((cnt, n), dtype=torch.float64, device=dev)
    BLK = 1000
    x0 = torch.randn(nx, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(33))
    dloc = torch.zeros(m, dtype=torch.float64, device=dev)   # D x0 on the D rows owned here
    # global row 0: [-1 0]; rows 1..md: [0 -D]; row md+1: zeros; rows md+2..: [0 -I]
    lo, hi = max(r0, 1), min(r0 + cnt, md + 1)      # D rows in this shard (global indices)
    b0 = ((lo - 1) // BLK) * BLK
    while lo < hi and b0 < hi - 1:
        g = torch.Generator(device=dev)
        g.manual_seed(3 * 1000003 + b0 // BLK)
        blk = torch.randn((BLK, nx), dtype=torch.float64, device=dev, generator=g) / np.sqrt(nx)
        s, e = max(b0, lo - 1), min(b0 + BLK, hi - 1)   # D-row indices (0-based inside D)
        if e > s:
            A[s + 1 - r0:e + 1 - r0, 1:] = -blk[s - b0:e - b0]
            dloc[s + 1:e + 1] = blk[s - b0:e - b0] @ x0
        b0 += BLK
    if r0 == 0 and cnt > 0:
        A[0, 0] = -1.0
    lo2, hi2 = max(r0, md + 2), min(r0 + cnt, m)
    if hi2 > lo2:
        gi = torch.arange(lo2, hi2, device=dev)
        A[gi - r0, 1 + (gi - (md + 2))] = -1.0
    if world > 1:
        dist.all_reduce(dloc)
    noise = torch.randn(md, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(34))
    dvec = dloc[1:md + 1] + 0.1 * noise
    rho = 0.5 * float(torch.linalg.norm(x0))
    b = np.concatenate([[0.0], -dvec.cpu().numpy(), [rho], np.zeros(nx)])
    c = np.zeros(n)
    c[0] = 1.0
    cones1, cones2 = [("SOC", md + 1), ("SOC", nx + 1)], [("Free", n)]
    H = fos.Handle(local)
    if world == 1:   # dense block of the D rows through K1, the -I rows as CSR + CSC (default: automatic)
        if getattr(args, "hybrid", False):
            H.set_option("hybrid_rows", 1)
        elif getattr(args, "no_hybrid", False):
            H.set_option("hybrid_rows", 0)
    if world > 1:
        cid = parallel.exchange_comm_id(rank, parallel.nccl_unique_id, dist)
        parallel.init_comm(H, rank, world, cid)
    t1, l1 = _cone_arrays(cones1, m, "constraint")
    t2, l2 = _cone_arrays(cones2, n, "variable")
    H.ck(H.L.fos_load_conic_dense(H.h, m, n, C.c_void_p(A.data_ptr()), n, 1, r0, cnt, _d(b), _d(c), len(t1), _i32p(t1),
                                  _i64p(l1), len(t2), _i32p(t2), _i64p(l2)))
    if world > 1 and args.exchange == "p2p":
        parallel.enable_p2p_exchange(H, rank, world, dist)
    H.set_algorithm(fos.GAPA())
    H.set_initial_iterate()
    H.ck(H.L.fos_begin_solve(H.h))
    stream = torch.cuda.ExternalStream(H.stream(), device=dev)
    W, K = args.warmup, args.iters
    H.run(1, W, 100, 1e-5)
    p0, cg0 = H.info("total_passes"), H.info("total_cg")
    if world > 1:
        dist.barrier()
    ms, (done, st, rec, _) = timed(torch, stream, lambda: H.run(W + 1, K, 100, 1e-5))
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    passes, cgs = H.info("total_passes") - p0, H.info("total_cg") - cg0
    peak, peak_src = peak_hbm()
    bytes_pass = 8.0 * m * n
    gbs = passes * bytes_pass / (ms / 1e3) / 1e9
    if rank == 0:
        print(json.dumps({"config": "C3", "algorithm": "GAPA()", "n_gpus": world, "exchange": args.exchange if world > 1 else None,
                          "m": m, "n": n, "matrix_gb": bytes_pass / 1e9, "storage_kind": int(H.info("storage_kind")),
                          "bytes_streamed_per_pass_gb": H.info("bytes_per_pass") * (world if world > 1 else 1) / 1e9,
                          "iterations_timed": int(done), "ms_per_iteration": ms / max(done, 1),
                          "iterations_per_s": done / (ms / 1e3), "cg_iterations_per_step": cgs / max(done, 1),
                          "passes_over_A_per_step": passes / max(done, 1), "aggregate_gbs": gbs,
                          "per_gpu_gbs": gbs / world, "peak_gbs_per_gpu": peak, "frac_per_gpu": gbs / world / peak,
                          "peak_source": peak_src, "status": int(st),
                          "last_p_d_g": [float(x) for x in rec[-1][1:4]] if len(rec) else None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4", "c5"])
    ap.add_argument("--nprob", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--md", type=int, default=100000)
    ap.add_argument("--nx", type=int, default=20000)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--batch-ctas", type=int, default=0, help="c5: persistent CTAs of the batch kernel (0 = default)")
    ap.add_argument("--dense-batch", action="store_true", help="c5: stream every row as dense FP64 (batch_hybrid = 0)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--hybrid", action="store_true", help="c3, one GPU: force the hybrid row storage (hybrid_rows = 1)")
    ap.add_argument("--no-hybrid", action="store_true", help="c3, one GPU: all rows as dense tiles (hybrid_rows = 0)")
    args = ap.parse_args()
    {"c3": run_c3, "c4": run_c4, "c5": run_c5}[args.config](args)


if __name__ == "__main__":
    main()
