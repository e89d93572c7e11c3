#!/usr/bin/env python
"""bench.py -- GAP/DR outer iterations per second on BASELINE.json's config 2:
"Dense LASSO via HSDE, A 20000x40000 FP64, DR(0.5) with CG affine projection, 1 B200".

A "step" is one outer DR iteration of the hot path (affine projection by CG on the KKT operator
= k+1 passes over A with the fused right-hand side, cone projection, relaxation).  The timed region is exactly
`--steps` consecutive iterations after `--warmup` untimed ones; A (6.4 GB) does not fit L2, so
every pass streams it from HBM ("inputs larger than L2", no flush needed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--m M --n N]

N > 1 (torchrun, one rank per GPU): the same matrix is row-sharded over the ranks and the A'
partial sums are all-reduced with NCCL inside the library ("scaling": "strong").
--impl reference times the CPU restatement of the reference (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "DR outer iterations/s (dense HSDE LASSO 20000x40000 FP64, CG affine projection)"
UNIT = "iterations/s"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons of one GPU, polled every 20 ms through NVML while the timed
    region runs (nvidia-smi -lms 200 is too coarse for a ~0.5 s region); nvidia-smi as fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.smax = None
        self.how = "nvml"

    def _loop_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        try:
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.smax = None
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def _loop_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.smax = float(f[1])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[2:6]):
                    if val.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass

    def start(self):
        try:
            import pynvml  # noqa: F401
            target = self._loop_nvml
        except Exception:
            target, self.how = self._loop_smi, "nvidia-smi"
        self.thread = threading.Thread(target=target, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=6)
        sm = self.samples
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(sm), "how": self.how}


# ------------------------------------------------------------------------------------------------
# problem construction: A = randn(m,n)/sqrt(n), K1 = Zero(m/2) + NonNeg(m/2),
# strictly feasible primal and dual by construction, variables free.
# ------------------------------------------------------------------------------------------------
ROW_BLOCK = 1000  # rows generated per RNG call; the matrix is identical for every GPU count


def gen_rows_device(torch, dev, r0, r1, n, seed):
    """Rows [r0, r1) of A on the device, block-seeded so that sharding does not change the matrix."""
    out = torch.empty((r1 - r0, n), dtype=torch.float64, device=dev)
    scale = 1.0 / np.sqrt(n)
    b = (r0 // ROW_BLOCK) * ROW_BLOCK
    while b < r1:
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 1000003 + b // ROW_BLOCK)
        blk = torch.randn((ROW_BLOCK, n), dtype=torch.float64, device=dev, generator=g)
        lo, hi = max(b, r0), min(b + ROW_BLOCK, r1)
        out[lo - r0:hi - r0] = blk[lo - b:hi - b] * scale
        b += ROW_BLOCK
    return out


def small_vectors(m, n, seed):
    rng = np.random.default_rng(seed)
    xi = rng.standard_normal(n)
    h = m // 2
    s = np.concatenate([np.zeros(h), np.abs(rng.standard_normal(m - h))])          # s* in K1
    y = np.concatenate([rng.standard_normal(h), np.abs(rng.standard_normal(m - h))])  # y* in K1*
    return xi, s, y, [("Zero", h), ("NonNeg", m - h)]


def cpu_reference_rate(m, n, steps, warmup, seed=2):
    """The restated reference (oracle/) timed on the host cores, on the NAMED configuration itself: the dense
    m x n matrix is stored the way the reference stores it (SparseMatrixCSC{Float64,Int64}, 16 B per entry,
    src/types.jl:35), every KKT product makes the reference's four CSC passes, CG as in conjugategradients.jl.
    The sparse products and CG sweeps run on all host threads (OpenMP build of the same C file,
    oracle/libfos_oracle_mt.so) -- the Julia original is single-threaded there, so this baseline is faster than
    the original would be.  Times `steps` consecutive DR iterations after `warmup` untimed ones."""
    from oracle import fos_oracle as fo
    variant = "mt"
    try:
        fo.build(variant="mt")
    except Exception:
        variant = ""    # no libgomp: single thread, exactly like the reference
        fo.build()
    t0 = time.perf_counter()
    O = fo.OracleConicDenseBig(m, n, seed, small_vectors, variant=variant)
    t_setup = time.perf_counter() - t0
    O.set_algorithm("GAP", 0.5, 2.0, 2.0)
    O.set_iterate(O.initial_value())
    if warmup > 0:
        O.run(1, warmup, checki=100, eps=1e-5)
    per_step, cg = [], []
    for k in range(steps):
        t0 = time.perf_counter()
        O.run(warmup + 1 + k, 1, checki=100, eps=1e-5)
        per_step.append(time.perf_counter() - t0)
        cg.append(int(O.cgiter))
    dt = float(sum(per_step))
    threads = fo.host_threads(variant)
    passes = sum(4 * k + 6 for k in cg)
    return {"value": steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"restated reference (C, CSC Float64/Int64 = {16 * m * n / 1e9:.1f} GB, 4 passes per KKT product) on "
                      f"the full {m}x{n} matrix, DR iterations {warmup + 1}..{warmup + steps} "
                      f"({sum(cg) / steps:.1f} CG iterations and {passes / steps:.0f} CSC passes per step), "
                      f"{threads} OpenMP thread(s) of {os.cpu_count()} host cores; the Julia original runs these loops on 1",
            "seconds": dt, "setup_seconds": t_setup, "cg_iterations_per_step": sum(cg) / steps,
            "host_gbs": passes * 16.0 * m * n / dt / 1e9}


def parity_block(fos, device):
    """Untimed correctness evidence printed with every bench line: a 1/10-scale twin of the workload
    (2000 x 4000, same recipe) in lock-step with the CPU oracle through the C ABI -- (a) well-conditioned
    scaling: the north-star 1e-10 bar, (b) the unscaled twin against the reference's arithmetic (C oracle) and
    against the exact (long-double reductions) restatement, see tests/test_gpu_exact.py."""
    from oracle import fos_oracle as fo
    from fos_b200 import problems
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import load_conic, rel_err, sync_state_from_oracle
    fo.build()
    fo.build(variant="hp")
    out = {"twin": "lasso_like(2000, 4000, seed=2): 1/10-scale twin of the workload, DR iterations 3..5 in lock-step from the C "
                   "oracle's state, S1 call counter advanced to 40 (CG tolerance 4e-5); exact = long-double reductions"}
    for tag, scale in (("well_conditioned", 0.02), ("unscaled", 1.0)):
        P = problems.lasso_like(2000, 4000, seed=2, scale=scale)
        O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
        X = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, variant="hp")
        H = load_conic(fos, P, storage="dense_direct")
        for o in (O, X):
            o.set_algorithm("GAP", 0.5, 2.0, 2.0)
        H.set_algorithm(fos.DR(0.5))
        O.set_iterate(O.initial_value())
        H.ck(H.L.fos_begin_solve(H.h))
        O.run(1, 1, checki=100000, eps=1e-12)
        O.set_scalar("s1_calls", 40)
        dev_c, dev_x, c_x, cg_match, cgs = 0.0, 0.0, 0.0, True, []
        for i in range(2, 6):
            sync_state_from_oracle(H, O, "DR")
            X.set_state("x", O.get_state("x"))
            X.set_state("xinit", O.get_state("xinit"))
            X.set_scalar("s1_calls", O.s1_calls)
            O.run(i, 1, checki=100000, eps=1e-12)
            X.run(i, 1, checki=100000, eps=1e-12)
            H.run(i, 1, 100000, 1e-12)
            cg_match = cg_match and int(H.info("cgiter")) == int(O.cgiter)
            cgs.append(int(O.cgiter))
            if i == 2:
                continue    # transient right after the tolerance jump (see tests/test_gpu_scale.py)
            z = H.get_iterate()
            dev_c = max(dev_c, rel_err(z, O.get_state("x")))
            dev_x = max(dev_x, rel_err(z, X.get_state("x")))
            c_x = max(c_x, rel_err(O.get_state("x"), X.get_state("x")))
        out[tag] = {"gpu_vs_oracle": dev_c, "gpu_vs_exact": dev_x, "oracle_vs_exact": c_x,
                    "cg_counts_match": bool(cg_match), "cg_iterations": cgs}
        del H
    wc, us = out["well_conditioned"], out["unscaled"]
    out["pass"] = bool(wc["gpu_vs_exact"] < 1e-10 and wc["gpu_vs_oracle"] < max(1e-10, 3 * wc["oracle_vs_exact"]) and
                       wc["cg_counts_match"] and us["gpu_vs_exact"] <= 3 * us["oracle_vs_exact"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=20000)
    ap.add_argument("--n", type=int, default=40000)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--matvec-impl", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip time-to-eps, the parity block and the N=1 replay")
    ap.add_argument("--tte-max-iters", type=int, default=2000, help="iteration cap of the time-to-eps solve")
    ap.add_argument("--tail-blocks", type=int, default=0, help="blocks of the fused CG-tail kernel (0 = one per SM)")
    ap.add_argument("--no-graphs", action="store_true", help="kernel-per-launch path with host synchronisation per CG batch")
    ap.add_argument("--k1-balance", type=int, default=-1, help="0 = even split of the tiles over the SMs, 1 = sized to the SMs' measured speed (default)")
    ap.add_argument("--tail-trace", action="store_true", help="phase timing of the fused CG tail (extra key tail_trace_us)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: fused peer-memory exchange kernel (default) or fold + ncclAllReduce")
    args = ap.parse_args()
    W = max(args.warmup, 0)
    K = max(args.steps, 1)
    m, n = args.m, args.n
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"C2 dense LASSO-shaped HSDE conic program, A {m}x{n} FP64 row-major in HBM "
                          f"({8 * m * n / 1e9:.1f} GB), K1 = Zero({m // 2}) + NonNeg({m - m // 2}), DR(0.5), "
                          f"CG affine projection with the reference's 0.2^sqrt(i) tolerance schedule",
              "m": m, "n": n, "algorithm": "DR(0.5)", "iterations_timed": f"{W + 1}..{W + K}",
              "l2": "inputs larger than L2 (A is streamed from HBM every pass; no flush needed)",
              "parallelism": "single GPU" if world == 1 else
              f"A row-sharded over {world} GPUs, " + ("fused peer-memory (NVLink, CUDA IPC) exchange kernel"
                                                      if args.exchange == "p2p" else "NCCL all-reduce")}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_reference_rate(m, n, K, W, args.seed)
        config["parallelism"] = f"host CPU, {cb['cores']} thread(s)"
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 * cb["seconds"] / K, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import fos_b200 as fos
    from fos_b200 import parallel
    from fos_b200.model import _cone_arrays, _d, _i32p, _i64p
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a GPU: the fos_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- build the problem on the device (setup, untimed) --------------------------------------
    r0, cnt = parallel.row_shard(m, rank, world)
    A_loc = gen_rows_device(torch, dev, r0, r0 + cnt, n, args.seed)
    xi, s, y, cones = small_vectors(m, n, args.seed)
    xi_d = torch.from_numpy(xi).to(dev)
    y_d = torch.from_numpy(y).to(dev)
    b_full = torch.zeros(m, dtype=torch.float64, device=dev)
    b_full[r0:r0 + cnt] = A_loc @ xi_d
    c_full = -(A_loc.T @ y_d[r0:r0 + cnt])
    if world > 1:
        dist.all_reduce(b_full)
        dist.all_reduce(c_full)
    b = b_full.cpu().numpy() + s
    c = c_full.cpu().numpy()
    del xi_d, y_d, b_full, c_full
    torch.cuda.synchronize()

    def apply_options(Hx):
        Hx.set_option("matvec_impl", args.matvec_impl)
        if args.tail_blocks:
            Hx.set_option("tail_blocks", args.tail_blocks)
        if args.no_graphs:
            Hx.set_option("use_graphs", 0)
        if args.k1_balance >= 0:
            Hx.set_option("k1_balance", args.k1_balance)

    H = fos.Handle(local_rank)
    apply_options(H)
    if world > 1:
        cid = parallel.exchange_comm_id(rank, parallel.nccl_unique_id, dist)
        parallel.init_comm(H, rank, world, cid)
    t1, l1 = _cone_arrays(cones, m, "constraint")
    t2, l2 = _cone_arrays([("Free", n)], n, "variable")
    H.ck(H.L.fos_load_conic_dense(H.h, m, n, C.c_void_p(A_loc.data_ptr()), n, 1, r0, cnt, _d(b), _d(c), len(t1),
                                  _i32p(t1), _i64p(l1), len(t2), _i32p(t2), _i64p(l2)))
    if world > 1 and args.exchange == "p2p":
        if not parallel.enable_p2p_exchange(H, rank, world, dist):
            args.exchange = "nccl"   # CUDA IPC unavailable on this box: every rank fell back together
            config["parallelism"] = f"A row-sharded over {world} GPUs, NCCL all-reduce (peer-memory mapping unavailable)"
    H.set_algorithm(fos.DR(0.5))
    H.set_initial_iterate()
    H.ck(H.L.fos_begin_solve(H.h))
    stream = torch.cuda.ExternalStream(H.stream(), device=dev)

    def stream_of(Hx):
        return torch.cuda.ExternalStream(Hx.stream(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_handle():
        """A fresh solver handle on the same device-resident matrix (shard), DR(0.5), initial iterate."""
        Hn = fos.Handle(local_rank)
        apply_options(Hn)
        if world > 1:
            cidn = parallel.exchange_comm_id(rank, parallel.nccl_unique_id, dist)
            parallel.init_comm(Hn, rank, world, cidn)
        Hn.ck(Hn.L.fos_load_conic_dense(Hn.h, m, n, C.c_void_p(A_loc.data_ptr()), n, 1, r0, cnt, _d(b), _d(c),
                                        len(t1), _i32p(t1), _i64p(l1), len(t2), _i32p(t2), _i64p(l2)))
        if world > 1 and args.exchange == "p2p":
            parallel.enable_p2p_exchange(Hn, rank, world, dist)
        Hn.set_algorithm(fos.DR(0.5))
        Hn.set_initial_iterate()
        Hn.ck(Hn.L.fos_begin_solve(Hn.h))
        return Hn

    # ---- warm-up -----------------------------------------------------------------------------------
    if W > 0:
        H.run(1, W, 100, 1e-5)
    launches0 = H.info("launches")
    cg0, passes0 = H.info("total_cg"), H.info("total_passes")

    # ---- timed region: K iterations, inputs resident in HBM -----------------------------------------
    # (the library's default path: one CUDA graph per outer iteration, no host synchronisation in between)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_wall0 = time.perf_counter()
    done, st, rec, _ = H.run(W + 1, K, 100, 1e-5)
    e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    assert done == K, f"solver stopped after {done} of {K} timed iterations (status {st}); lower --steps"
    launches = H.info("launches") - launches0
    cg_iters = H.info("total_cg") - cg0
    passes = H.info("total_passes") - passes0   # executed passes over A (predicated no-op launches excluded)
    bytes_pass = H.info("bytes_per_pass")
    value = K / (ms_total / 1e3)

    # ---- roofline pass: the SAME iterations W+1..W+K on a fresh handle, every launch of the dominant kernel
    # bracketed by CUDA events on the library's stream ("profile_matvec": kernel-per-launch path, since events
    # cannot be recorded inside a graph's WHILE body) ------------------------------------------------------
    Hp = make_handle()
    if W > 0:
        Hp.run(1, W, 100, 1e-5)
    Hp.set_option("profile_matvec", 1)
    if args.tail_trace:
        Hp.set_option("tail_trace", 1)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream_of(Hp))
    Hp.run(W + 1, K, 100, 1e-5)
    p1.record(stream_of(Hp))
    barrier()
    prof_ms_total = p0.elapsed_time(p1)
    mv2_ms, mv2_n = Hp.info("mv2_ms"), Hp.info("mv2_n")
    mv1_ms, mv1_n = Hp.info("mv1_ms"), Hp.info("mv1_n")
    tail_ms, tail_n = Hp.info("tail_ms"), Hp.info("tail_n")
    Hp.set_option("profile_matvec", 0)
    tail_trace = None
    if args.tail_trace:
        tt_ = Hp.tail_trace()
        mhz = clocks.get("sm_mhz") or 1900.0
        tail_trace = {"phases": ["fold+push", "-", "-", "poll+Ap", "allreduce1", "update", "allreduce2", "dir"],
                      "us_per_launch": [[float(tt_[b, k] / max(tt_[b, 15], 1) / mhz) for k in range(8)] for b in range(3)],
                      "launches": float(tt_[0, 15]), "sm_mhz": mhz, "rank": rank}
        if rank != 0:
            sys.stderr.write("rank %d tail_trace %s\n" % (rank, json.dumps(tail_trace)))
    del Hp

    # ---- e2e: the SAME iterations W+1..W+K through the C ABI with HOST buffers ---------------------
    # A second handle on the same device matrix replays the solve; every step moves the iterate in
    # from host memory (H2D), runs one iteration with a residual check and reads the iterate and the
    # p/d/g record back (D2H), all inside the timed region.
    H2 = fos.Handle(local_rank)
    apply_options(H2)
    if world > 1:
        cid2 = parallel.exchange_comm_id(rank, parallel.nccl_unique_id, dist)
        parallel.init_comm(H2, rank, world, cid2)
    H2.ck(H2.L.fos_load_conic_dense(H2.h, m, n, C.c_void_p(A_loc.data_ptr()), n, 1, r0, cnt, _d(b), _d(c), len(t1),
                                    _i32p(t1), _i64p(l1), len(t2), _i32p(t2), _i64p(l2)))
    if world > 1 and args.exchange == "p2p":
        parallel.enable_p2p_exchange(H2, rank, world, dist)
    H2.set_algorithm(fos.DR(0.5))
    H2.set_initial_iterate()
    H2.ck(H2.L.fos_begin_solve(H2.h))
    if W > 0:
        H2.run(1, W, 100, 1e-5)
    z = H2.get_iterate()
    N = z.size
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        H2.set_iterate(z)
        H2.run(W + 1 + k, 1, 100, 1e-5)    # same check interval as the timed region
        z = H2.get_iterate()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e = {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(N * 8),
           "d2h_bytes_per_step": int(N * 8),
           "note": "per step: fos_set_iterate (host->device) + fos_run(1 iteration) + fos_get_iterate "
                   "(device->host) on a second handle; same iterations and check interval as the timed region"}
    del H2

    # ---- time to eps = 1e-5 (BASELINE.json's metric, second half): solve!(model) from the initial iterate -----
    # through fos_solve (solverwrapper.jl:2-17: iterations with a check every 100, getsol, final check)
    tte = None
    if not args.no_extras:
        H3 = make_handle()
        z0 = H3.get_iterate()
        barrier()
        t0 = time.perf_counter()
        done3, st3, rec3, _ = H3.solve(args.tte_max_iters, 100, 1e-5)
        torch.cuda.synchronize()
        tte_s = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([tte_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tte_s = float(tt.item())
        tte = {"seconds": tte_s, "eps": 1e-5, "iterations": int(done3), "status": fos.model.STATUS_SYMBOLS[st3],
               "checki": 100, "max_iters": args.tte_max_iters, "cg_iterations_total": int(H3.info("total_cg")),
               "passes_over_A": int(H3.info("total_passes")),
               "last_record": {"i": int(rec3[-1, 0]), "p": float(rec3[-1, 1]), "d": float(rec3[-1, 2]),
                               "g": float(rec3[-1, 3])} if len(rec3) else None,
               "hbm_gbs": bytes_pass * H3.info("total_passes") / tte_s / 1e9,
               "note": "wall clock of fos_solve from z0 (tau = kappa = 1), max over ranks; includes getsol and the "
                       "final check"}
        del z0

    # ---- N > 1: the sharded path against one GPU holding the whole matrix (rank 0) ----------------------
    #   free-running: iterations 1..W+K on N shards vs the same iterations on one GPU;
    #   lock-step:    from the state the sharded time-to-eps solve ended in (tau > 0, residuals finite), ONE more
    #                 iteration with a residual check on both -- iterate, p/d/g record and CG count side by side
    multi = None
    if world > 1 and not args.no_extras:
        import zlib
        z_sh = H.get_iterate()
        crc = torch.tensor([float(zlib.crc32(z_sh.tobytes()))], dtype=torch.float64, device=dev)
        crcs = [torch.empty_like(crc) for _ in range(world)]
        dist.all_gather(crcs, crc)
        x3, xi3, s13 = H3.get_state("x"), H3.get_state("xinit"), H3.info("s1_calls")
        i3 = int(done3) + 1
        _, _, rec_sh, _ = H3.run(i3, 1, 1, 1e-5)
        z3 = H3.get_iterate()
        cg_sh = int(H3.info("cgiter"))
        if rank == 0:
            A_full = gen_rows_device(torch, dev, 0, m, n, args.seed)
            Hc = fos.Handle(local_rank)
            apply_options(Hc)
            Hc.ck(Hc.L.fos_load_conic_dense(Hc.h, m, n, C.c_void_p(A_full.data_ptr()), n, 1, 0, m, _d(b), _d(c),
                                            len(t1), _i32p(t1), _i64p(l1), len(t2), _i32p(t2), _i64p(l2)))
            Hc.set_algorithm(fos.DR(0.5))
            Hc.set_initial_iterate()
            Hc.ck(Hc.L.fos_begin_solve(Hc.h))
            Hc.run(1, W + K, 100, 1e-5)
            z_1 = Hc.get_iterate()
            cg_free_1 = int(Hc.info("total_cg"))
            Hc.set_state("x", x3)
            Hc.set_state("xinit", xi3)
            Hc.set_info("s1_calls", s13)
            _, _, rec_1, _ = Hc.run(i3, 1, 1, 1e-5)
            zc = Hc.get_iterate()
            rel = lambda u, v: float(np.abs(u - v).max() / max(np.abs(v).max(), 1e-300))
            recd = lambda r: {"i": int(r[0, 0]), "p": float(r[0, 1]), "d": float(r[0, 2]), "g": float(r[0, 3]),
                              "cgiter": int(r[0, 8]), "status": int(r[0, 9])}
            multi = {"ranks_bitwise_identical": bool(all(float(t.item()) == float(crc.item()) for t in crcs)),
                     "iterate_crc32": int(crc.item()),
                     "free_running": {"iterations": f"1..{W + K}", "iterate_deviation_vs_n1": rel(z_sh, z_1),
                                      "cg_iterations_sharded": int(cg0 + cg_iters), "cg_iterations_n1": cg_free_1},
                     "lockstep": {"iteration": i3, "iterate_deviation_vs_n1": rel(z3, zc),
                                  "record_sharded": recd(rec_sh), "record_n1": recd(rec_1),
                                  "cg_iterations_sharded": cg_sh, "cg_iterations_n1": int(Hc.info("cgiter"))},
                     "note": "row shards associate the A' sums differently from one GPU; free-running, the truncated CG "
                             "amplifies that (DESIGN.md parity budget); in lock-step one iteration agrees to rounding"}
            del Hc, A_full
        barrier()
    if tte is not None:
        del H3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    ach = (bytes_pass * mv2_n / (mv2_ms / 1e3)) / 1e9 if mv2_ms > 0 else None
    # DRAM bytes per launch from the committed ncu --set full capture: valid only for the shape it was taken on
    # (the single-GPU 20000 x 40000 pass); null for any other shard
    traffic = None
    tf = ROOT / "profiles" / "k1_traffic.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            if abs(float(tj.get("algorithmic_bytes_per_launch", -1)) - float(bytes_pass)) < 1:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k1_dual_matvec_tma<2> (fused A*[x1 x2] and A'*[y1 y2], one pass over A)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_pass, "launches_timed": int(mv2_n),
                "avg_launch_ms": (mv2_ms / mv2_n) if mv2_n else None,
                "share_of_step": ((mv2_ms + mv1_ms) / prof_ms_total) if prof_ms_total > 0 else None,
                "measured_in": "event-bracketed replay of the timed iterations (kernel-per-launch path): "
                               f"{prof_ms_total / K:.3f} ms per step there",
                "sm_balanced": bool(H.info("k1_balanced")),
                "sm_time_spread_before_after": [H.info("k1_spread_before"), H.info("k1_spread_after")],
                "whole_iteration_gbs": bytes_pass * passes / (ms_total / 1e3) / 1e9,
                "whole_iteration_frac": bytes_pass * passes / (ms_total / 1e3) / 1e9 / peak}
    parity = None
    if not args.no_extras:
        try:
            parity = parity_block(fos, local_rank)
        except Exception as ex:   # evidence, not the measurement: report, do not abort the bench line
            parity = {"error": repr(ex)}
    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_rate(m, n, 3, W, args.seed)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cb,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "cg_tail_avg_launch_us": (1e3 * tail_ms / tail_n) if tail_n else None,
            "cg_iterations_per_step": cg_iters / K, "passes_over_A_per_step": passes / K,
            "wall_ms_per_step": t_wall * 1e3 / K, "status_after_timed": int(st),
            "iteration_path": "kernel per launch" if args.no_graphs else "CUDA graph per outer iteration (CG loop = WHILE node)"}
    line["time_to_eps"] = tte
    line["parity"] = parity
    if multi is not None:
        line["multi_gpu_parity"] = multi
    if tail_trace is not None:
        line["tail_trace_us"] = tail_trace
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
