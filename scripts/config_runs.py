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
    import torch
    import fos_b200 as fos
    dev = torch.device("cuda", 0)
    B, rows, cols = args.nprob, 256, 512
    m, n = rows + 1 + cols, cols + 1
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    D = torch.randn((B, rows, cols), dtype=torch.float64, device=dev, generator=g) / np.sqrt(cols)
    d = torch.randn((B, rows), dtype=torch.float64, device=dev, generator=g)
    A = torch.zeros((B, m, n), dtype=torch.float64, device=dev)
    A[:, 0, 0] = -1.0
    A[:, 1:rows + 1, 1:] = -D
    idx = torch.arange(cols, device=dev)
    A[:, rows + 1 + idx, 1 + idx] = -1.0
    b = np.zeros((B, m))
    b[:, 1:rows + 1] = -d.cpu().numpy()
    c = np.zeros((B, n))
    c[:, 0] = 1.0
    cones1, cones2 = [("SOC", rows + 1), ("NonNeg", cols)], [("Free", n)]
    peak, peak_src = peak_hbm()
    for name, alg in (("FISTA", fos.FISTA()), ("Dykstra", fos.Dykstra())):
        H = fos.Handle(0)
        H.load_conic_batch((B, m, n), b, c, cones1, cones2, device_ptr=(A.data_ptr(), n, m * n))
        H.set_algorithm(alg)
        stream = torch.cuda.ExternalStream(H.stream(), device=dev)
        H.ck(H.L.fos_begin_solve_batch(H.h))
        W, K = args.warmup, args.iters
        H.run_batch(1, W, 10 ** 9, 1e-5)
        p0 = H.info_batch("total_passes").sum()
        cg0 = H.info_batch("total_cg").sum()
        ms, (done, st, recs) = timed(torch, stream, lambda: H.run_batch(W + 1, K, K, 1e-5))
        passes = H.info_batch("total_passes").sum() - p0
        cgs = H.info_batch("total_cg").sum() - cg0
        bytes_pass = 8.0 * m * n
        gbs = passes * bytes_pass / (ms / 1e3) / 1e9
        line = {"config": "C5", "algorithm": name, "nprob": B, "m": m, "n": n, "iterations_timed": K,
                "ms_total": ms, "problem_iterations_per_s": B * K / (ms / 1e3),
                "iterations_per_s_per_problem_stream": K / (ms / 1e3),
                "cg_iterations_per_step": cgs / (B * K), "passes_over_A_per_step": passes / (B * K),
                "algorithmic_bytes_per_pass": bytes_pass, "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak,
                "peak_source": peak_src, "all_iterated": bool((done == K).all()),
                "check_p_median": float(np.median([r[-1][1] for r in recs if len(r)])) if K else None}
        if args.cpu:
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
        print(json.dumps(line), flush=True)
        del H


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
            "cg_iterations_per_step": cgs / max(done, 1), "psd_projection_2x_ms": ms_psd2, "jacobi_sweeps": sweeps,
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
    import ctypes as C
    import torch
    import fos_b200 as fos
    from fos_b200.model import _cone_arrays, _d, _i32p, _i64p
    dev = torch.device("cuda", 0)
    md, nx = args.md, args.nx
    m, n = md + 1 + nx + 1, nx + 1
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    A = torch.zeros((m, n), dtype=torch.float64, device=dev)
    A[0, 0] = -1.0
    blk = 10000
    for r0 in range(0, md, blk):
        r1 = min(md, r0 + blk)
        A[1 + r0:1 + r1, 1:] = -torch.randn((r1 - r0, nx), dtype=torch.float64, device=dev, generator=g) / np.sqrt(nx)
    idx = torch.arange(nx, device=dev)
    A[md + 2 + idx, 1 + idx] = -1.0
    x0 = torch.randn(nx, dtype=torch.float64, device=dev, generator=g)
    dvec = (-A[1:md + 1, 1:]) @ x0 + 0.1 * torch.randn(md, dtype=torch.float64, device=dev, generator=g)
    rho = 0.5 * float(torch.linalg.norm(x0))
    b = np.concatenate([[0.0], -dvec.cpu().numpy(), [rho], np.zeros(nx)])
    c = np.zeros(n)
    c[0] = 1.0
    cones1, cones2 = [("SOC", md + 1), ("SOC", nx + 1)], [("Free", n)]
    H = fos.Handle(0)
    t1, l1 = _cone_arrays(cones1, m, "constraint")
    t2, l2 = _cone_arrays(cones2, n, "variable")
    H.ck(H.L.fos_load_conic_dense(H.h, m, n, C.c_void_p(A.data_ptr()), n, 1, 0, m, _d(b), _d(c), len(t1), _i32p(t1),
                                  _i64p(l1), len(t2), _i32p(t2), _i64p(l2)))
    H.set_algorithm(fos.GAPA())
    H.set_initial_iterate()
    H.ck(H.L.fos_begin_solve(H.h))
    stream = torch.cuda.ExternalStream(H.stream(), device=dev)
    W, K = args.warmup, args.iters
    H.run(1, W, 100, 1e-5)
    p0, cg0 = H.info("total_passes"), H.info("total_cg")
    ms, (done, st, rec, _) = timed(torch, stream, lambda: H.run(W + 1, K, 100, 1e-5))
    passes, cgs = H.info("total_passes") - p0, H.info("total_cg") - cg0
    peak, peak_src = peak_hbm()
    bytes_pass = H.info("bytes_per_pass")
    gbs = passes * bytes_pass / (ms / 1e3) / 1e9
    print(json.dumps({"config": "C3", "algorithm": "GAPA()", "m": m, "n": n, "matrix_gb": bytes_pass / 1e9,
                      "iterations_timed": int(done), "ms_per_iteration": ms / max(done, 1),
                      "iterations_per_s": done / (ms / 1e3), "cg_iterations_per_step": cgs / max(done, 1),
                      "passes_over_A_per_step": passes / max(done, 1), "achieved_gbs": gbs, "peak_gbs": peak,
                      "frac": gbs / peak, "peak_source": peak_src, "status": int(st)}), flush=True)


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
    args = ap.parse_args()
    {"c3": run_c3, "c4": run_c4, "c5": run_c5}[args.config](args)


if __name__ == "__main__":
    main()
