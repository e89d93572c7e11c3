"""Replays tests/test_gpu_units.py::test_psd_large_batched_cones several times and prints every stage's error."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import fos_b200 as fos  # noqa: E402
from fos_b200 import problems  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402

reps = int(sys.argv[1])
cases = [(512, 2), (129, 5), (640, 1), (1024, 1)]
refs = {}
for rep in range(reps):
    for d, nc in cases:
        H = fos.Handle(0)
        rng = np.random.default_rng(d + nc)
        plen = d * (d + 1) // 2
        X = rng.standard_normal((nc, plen))
        if nc >= 2:
            U = rng.standard_normal((d, 3))
            X[1] = problems.svec(U @ U.T - 0.2 * np.eye(d)) + 1e-6 * rng.standard_normal(plen)
        if nc >= 3:
            G = rng.standard_normal((d, d))
            X[2] = problems.svec(-(G @ G.T) - np.eye(d))
        if nc >= 4:
            X[3] = 0.0
        if (d, nc) not in refs:
            refs[(d, nc)] = np.stack([npo.prox_cone("SDP", X[k]) for k in range(nc)])
        ref = refs[(d, nc)]
        Y, ms, sw = H.time_psd(X, reps=1)
        e1 = [float(np.abs(Y[k] - ref[k]).max() / max(np.abs(ref[k]).max(), np.abs(X[k]).max())) for k in range(nc)]
        Y2, _, sw2 = H.time_psd(Y, reps=1)
        e2 = [float(np.abs(Y2[k] - Y[k]).max() / max(np.abs(Y).max(), 1.0)) for k in range(nc)]
        Yn, _, sw3 = H.time_psd(-X, reps=1)
        e3 = [float(np.abs(Y[k] - Yn[k] - X[k]).max() / np.abs(X).max()) for k in range(nc)]
        bad = max(e1) > 1e-12 or max(e2) > 1e-12 or max(e3) > 1e-12
        print(f"rep {rep} d={d} nc={nc} sweeps {sw},{sw2},{sw3} proj {max(e1):.1e} idem {['%.1e' % e for e in e2]} "
              f"moreau {max(e3):.1e} {'<<<<<< BAD' if bad else ''}", flush=True)
