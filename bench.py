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


def cpu_reference_rate(m_full, n_full, steps, warmup, seed=2, sample_div=10):
    """The restated reference (oracle/, single thread like the original's mat-vecs) on a bounded
    sample: the same recipe at (m/10) x (n/10) (1/100 of the elements), iterations 1..warmup+steps.
    One CSC pass costs time proportional to nnz, so iterations/s at full size = sample rate *
    (nnz_sample / nnz_full)."""
    from oracle import fos_oracle as fo
    fo.build()
    ms, ns = max(m_full // sample_div, 16), max(n_full // sample_div, 16)
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((ms, ns)) / np.sqrt(ns)
    xi, s, y, cones = small_vectors(ms, ns, seed)
    b = A @ xi + s
    c = -(A.T @ y)
    O = fo.OracleConic(c, A, b, cones, [("Free", ns)])
    O.set_algorithm("GAP", 0.5, 2.0, 2.0)
    O.set_iterate(O.initial_value())
    if warmup > 0:
        O.run(1, warmup, checki=100, eps=1e-5)
    t0 = time.perf_counter()
    O.run(warmup + 1, steps, checki=100, eps=1e-5)
    dt = time.perf_counter() - t0
    rate_sample = steps / dt
    scale = (ms * ns) / float(m_full * n_full)
    return {"value": rate_sample * scale, "unit": UNIT, "cores": fo.host_threads(), "kind": "port",
            "sample": f"restated reference (C, CSC, 4 passes per KKT product) on {ms}x{ns} = 1/{int(round(1/scale))} "
                      f"of the elements, same recipe, iterations {warmup + 1}..{warmup + steps}: "
                      f"{rate_sample:.3f} it/s, scaled by nnz ratio; host has {os.cpu_count()} cores, "
                      f"OPENBLAS_NUM_THREADS={os.environ.get('OPENBLAS_NUM_THREADS', 'unset')}",
            "seconds": dt}


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
    ap.add_argument("--tail-flags", type=int, default=-1, help="N > 1: hand-shake of the fused CG tail (0 block to block, 1 per rank)")
    ap.add_argument("--tail-blocks", type=int, default=0, help="blocks of the fused CG-tail kernel (0 = one per SM)")
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
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
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

    H = fos.Handle(local_rank)
    H.set_option("matvec_impl", args.matvec_impl)
    if args.tail_blocks:
        H.set_option("tail_blocks", args.tail_blocks)
    if args.tail_flags >= 0:
        H.set_option("tail_flags", args.tail_flags)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -----------------------------------------------------------------------------------
    if W > 0:
        H.run(1, W, 100, 1e-5)
    H.set_option("profile_matvec", 1)
    launches0 = H.info("launches")
    cg0, passes0 = H.info("total_cg"), H.info("total_passes")

    # ---- timed region: K iterations, inputs resident in HBM -----------------------------------------
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
    mv2_ms, mv2_n = H.info("mv2_ms"), H.info("mv2_n")
    mv1_ms, mv1_n = H.info("mv1_ms"), H.info("mv1_n")
    bytes_pass = H.info("bytes_per_pass")
    tail_ms, tail_n = H.info("tail_ms"), H.info("tail_n")
    H.set_option("profile_matvec", 0)
    value = K / (ms_total / 1e3)

    # ---- e2e: the SAME iterations W+1..W+K through the C ABI with HOST buffers ---------------------
    # A second handle on the same device matrix replays the solve; every step moves the iterate in
    # from host memory (H2D), runs one iteration with a residual check and reads the iterate and the
    # p/d/g record back (D2H), all inside the timed region.
    H2 = fos.Handle(local_rank)
    H2.set_option("matvec_impl", args.matvec_impl)
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    ach = (bytes_pass * mv2_n / (mv2_ms / 1e3)) / 1e9 if mv2_ms > 0 else None
    traffic = None
    tf = ROOT / "profiles" / "k1_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k1_dual_matvec_tma<2> (fused A*[x1 x2] and A'*[y1 y2], one pass over A)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_pass, "launches_timed": int(mv2_n),
                "avg_launch_ms": (mv2_ms / mv2_n) if mv2_n else None,
                "share_of_step": ((mv2_ms + mv1_ms) / ms_total) if ms_total > 0 else None,
                "whole_iteration_gbs": bytes_pass * passes / (ms_total / 1e3) / 1e9,
                "whole_iteration_frac": bytes_pass * passes / (ms_total / 1e3) / 1e9 / peak}
    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_rate(m, n, min(K, 40), W, args.seed)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cb,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "cg_tail_avg_launch_us": (1e3 * tail_ms / tail_n) if tail_n else None,
            "cg_iterations_per_step": cg_iters / K, "passes_over_A_per_step": passes / K,
            "wall_ms_per_step": t_wall * 1e3 / K, "status_after_timed": int(st)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
