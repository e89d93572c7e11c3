"""Diagnostics for the large-PSD kernel: repeated projections of one d x d matrix, error against LAPACK."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import fos_b200 as fos  # noqa: E402
from fos_b200 import problems  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402

for spec in sys.argv[1:]:
    d, seed = (int(x) for x in spec.split(":"))
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((1, d * (d + 1) // 2))
    ref = npo.prox_cone("SDP", X[0])
    for rep in range(4):
        H = fos.Handle(0)
        Y, ms, sweeps = H.time_psd(X, reps=1)
        err = np.abs(Y[0] - ref).max() / np.abs(ref).max()
        print(f"d={d} seed {seed} rep {rep}: {ms:.3f} ms sweeps {sweeps} err {err:.2e}", flush=True)
