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
    A = torch.from_numpy(Ms).cuda()
    torch.linalg.eigh(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        w, V = torch.linalg.eigh(A)
        P = (V * w.clamp_min(0).unsqueeze(-2)) @ V.transpose(-1, -2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, P.cpu().numpy()


def main():
    H = fos.Handle(0)
    rng = np.random.default_rng(0)
    cases = [(512, 2), (512, 1), (256, 2), (128, 8), (1024, 1), (64, 64), (16, 1024)]
    for d, nc in cases:
        Ms = np.zeros((nc, d, d))
        X = np.zeros((nc, d * (d + 1) // 2))
        for k in range(nc):
            G = rng.standard_normal((d, d))
            Ms[k] = (G + G.T) / 2
            X[k] = problems.svec(Ms[k])
        Y, ms, sweeps = H.time_psd(X, reps=5)
        ms_lib, Pl = lib_bar(Ms, 5)
        err = max(np.abs(problems.smat(Y[k]) - Pl[k]).max() / np.abs(Pl[k]).max() for k in range(nc))
        print(json.dumps({"d": d, "ncones": nc, "fos_ms": ms, "sweeps": sweeps, "cusolver_eigh_ms": ms_lib,
                          "speedup": ms_lib / ms, "max_rel_diff_vs_lib": err}), flush=True)


if __name__ == "__main__":
    main()
