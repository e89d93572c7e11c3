"""First-contact probe for the GPU box: exercises each device path once with loud diagnostics."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import fos_b200 as fos  # noqa: E402
from fos_b200 import problems  # noqa: E402
from helpers import load_conic, rel_err  # noqa: E402
from oracle import fos_oracle as fo  # noqa: E402

small = "--small" in sys.argv
shapes = [(91, 51), (700, 4500)] if small else [(91, 51), (17, 513), (700, 4500), (2600, 2100)]
for (m, n) in shapes:
    P = problems.lasso_like(m, n, seed=2)
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    v = np.random.default_rng(0).standard_normal(2 * (m + n + 1))
    ref = O.kkt_mul(v)
    for path, kw in (("plain", dict(storage="dense", matvec_impl=1)), ("sparse", dict(storage="sparse")),
                     ("tma", dict(storage="dense", matvec_impl=0))):
        t = time.time()
        H = load_conic(fos, P, **kw)
        y = H.kkt_mul(v)
        print(f"{m}x{n} {path:6s} kkt_mul rel err {rel_err(y, ref):.2e}  ({time.time() - t:.2f}s)", flush=True)
if not small:
    P = problems.lasso_like(4096, 8192, seed=3)
    H = load_conic(fos, P, storage="dense")
    for nv in (1, 2):
        ms, by = H.time_matvec(nv, 20)
        print(f"4096x8192 NV={nv}: {ms:.4f} ms/launch, {by / ms / 1e6:.1f} GB/s", flush=True)
    H.set_algorithm(fos.DR(0.5))
    t = time.time()
    done, st, rec, _ = H.run(1, 20, 10, 1e-5)
    print("20 DR iterations:", time.time() - t, "s; records:", rec[:, [0, 1, 2, 3, 8]].tolist(), flush=True)
print("probe done")
