"""K5 timing probe: PSD projection of ncones packed d x d matrices, fos_b200 (Jacobi kernels) vs the
library bar (torch.linalg.eigh = cuSOLVER syevd on the same GPU, plus the two GEMM-shaped steps).
Writes one JSON line per case."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import fos_b200 as fos  # noqa: E402
from fos_b200 import problems  # noqa: E402


def lib_bar(Ms, reps):
    """The library bar, two ways: ONE batched torch.linalg.eigh call (torch picks syevjBatched / a syevd loop), and
    one eigh (cuSOLVER syevd) call PER MATRIX back to back -- the fair comparator for a handful of large cones.
    Returns (best of the two in ms, batched ms, sequential ms, projections)."""
    A = torch.from_numpy(Ms).cuda()
    torch.linalg.eigh(A)
    torch.linalg.eigh(A[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        w, V = torch.linalg.eigh(A)
        P = (V * w.clamp_min(0).unsqueeze(-2)) @ V.transpose(-1, -2)
    e1.record()
    torch.cuda.synchronize()
    ms_batched = e0.elapsed_time(e1) / reps
    ms_seq = float("inf")
    if A.shape[0] <= 16:
        e0.record()
        for _ in range(reps):
            for k in range(A.shape[0]):
                wk, Vk = torch.linalg.eigh(A[k])
                Pk = (Vk * wk.clamp_min(0).unsqueeze(-2)) @ Vk.transpose(-1, -2)
        e1.record()
        torch.cuda.synchronize()
        ms_seq = e0.elapsed_time(e1) / reps
    return min(ms_batched, ms_seq), ms_batched, ms_seq, P.cpu().numpy()


def main():
    H = fos.Handle(0)
    rng = np.random.default_rng(0)
    cases = [(512, 2), (512, 1), (256, 2), (128, 8), (1024, 1), (64, 64), (32, 256), (24, 512), (16, 1024), (8, 2048), (3, 4096)]
    cases += [(48, 128), (96, 32)]
    if "--large" in sys.argv:
        cases = [(512, 2), (512, 1), (256, 2), (256, 1), (128, 8), (1024, 1), (768, 1), (384, 1)]
    for d, nc in cases:
        Ms = np.zeros((nc, d, d))
        X = np.zeros((nc, d * (d + 1) // 2))
        for k in range(nc):
            G = rng.standard_normal((d, d))
            Ms[k] = (G + G.T) / 2
            X[k] = problems.svec(Ms[k])
        Y, ms, sweeps = H.time_psd(X, reps=5)
        ms_lib, ms_b, ms_s, Pl = lib_bar(Ms, 5)
        err = max(np.abs(problems.smat(Y[k]) - Pl[k]).max() / np.abs(Pl[k]).max() for k in range(nc))
        print(json.dumps({"d": d, "ncones": nc, "fos_cold_ms": ms, "sweeps": sweeps, "cusolver_eigh_ms": ms_lib,
                          "cusolver_batched_call_ms": ms_b, "cusolver_sequential_calls_ms": None if ms_s == float("inf") else ms_s,
                          "speedup_vs_best_library": ms_lib / ms, "max_rel_diff_vs_lib": err}), flush=True)


if __name__ == "__main__":
    main()
