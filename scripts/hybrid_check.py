"""GPU check of the opt-in hybrid row storage of the single-problem path (option "hybrid_rows" = 1,
MatOp::init_hybrid in csrc/matop.cu): the contiguous block that holds every dense row is streamed by K1,
the remaining non-empty rows travel as CSR + CSC, one fold merges both.

NOT YET RUN ON A GPU (written after the round's GPU minutes were spent): run this first next round,
then turn the checks into tests/test_gpu_units.py cases and flip the default.

  python scripts/hybrid_check.py            # parity against the oracle + the all-dense path, then timing
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import fos_b200 as fos  # noqa: E402
from fos_b200 import problems  # noqa: E402
from helpers import ALG_SETUPS, load_conic, rel_err, set_alg_both, sync_state_from_oracle  # noqa: E402
from oracle import fos_oracle  # noqa: E402


def operators(P, label):
    O = fos_oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    Hd = load_conic(fos, P, storage="dense")
    Hh = load_conic(fos, P, storage="dense", hybrid_rows=1)
    kind, dr, sr = Hh.info("storage_kind"), Hh.info("hybrid_dense_rows"), Hh.info("hybrid_sparse_rows")
    rng = np.random.default_rng(0)
    N = 2 * (P.m + P.n + 1)
    worst = 0.0
    for _ in range(3):
        z = rng.standard_normal(N)
        for name in ("kkt_mul", "affine_prox"):
            a, b, c = getattr(Hh, name)(z), getattr(Hd, name)(z), getattr(O, name)(z)
            worst = max(worst, rel_err(a, c))
            assert rel_err(a, c) < (1e-10 if name == "affine_prox" else 1e-11) and rel_err(a, b) < 1e-11, (label, name, rel_err(a, c), rel_err(a, b))
        x, y = rng.standard_normal(P.n), rng.standard_normal(P.m)
        assert rel_err(Hh.a_mul(x, P.m, P.n), O.a_mul(x)) < 1e-12
        assert rel_err(Hh.a_mul(y, P.m, P.n, transpose=True), O.a_mul(y, transpose=True)) < 1e-12
    print(json.dumps({"check": "operators", "problem": label, "storage_kind": kind, "dense_rows": dr, "sparse_rows": sr,
                      "bytes_per_pass_hybrid": Hh.info("bytes_per_pass"), "bytes_per_pass_dense": Hd.info("bytes_per_pass"),
                      "worst_vs_oracle": worst}))
    return kind


def lockstep(P, alg, label, n_iter=30):
    O = fos_oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    H = load_conic(fos, P, storage="dense", hybrid_rows=1)
    set_alg_both(fos, H, O, alg)
    O.set_iterate(O.initial_value())
    H.ck(H.L.fos_begin_solve(H.h))
    worst = 0.0
    for i in range(1, n_iter + 1):
        sync_state_from_oracle(H, O, alg)
        O.run(i, 1, checki=5, eps=1e-12)
        H.run(i, 1, 5, 1e-12)
        assert H.info("cgiter") == O.cgiter, (label, alg, i)
        worst = max(worst, rel_err(H.get_iterate(), O.get_state("x")))
    assert worst < 1e-10, (label, alg, worst)
    print(json.dumps({"check": "lockstep", "problem": label, "alg": alg, "worst": worst}))


def timing(md, nx):
    import torch
    P = None
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    m, n = md + nx + 2, nx + 1
    A = torch.zeros((m, n), dtype=torch.float64, device=dev)
    A[0, 0] = -1.0
    A[1:md + 1, 1:] = -torch.randn((md, nx), dtype=torch.float64, device=dev, generator=g) / np.sqrt(nx)
    A[md + 2:, 1:] = -torch.eye(nx, dtype=torch.float64, device=dev)
    b = np.concatenate([[0.0], np.random.default_rng(1).standard_normal(md), [1.0], np.zeros(nx)])
    c = np.zeros(n)
    c[0] = 1.0
    torch.cuda.synchronize()
    out = {}
    for hyb in (0, 1):
        H = fos.Handle(0)
        H.set_option("hybrid_rows", hyb)
        from fos_b200 import model as M
        t1, l1 = M._cone_arrays([("SOC", md + 1), ("SOC", nx + 1)], m, "constraint")
        t2, l2 = M._cone_arrays([("Free", n)], n, "variable")
        import ctypes as C
        H.ck(H.L.fos_load_conic_dense(H.h, m, n, C.c_void_p(A.data_ptr()), n, 1, 0, m, M._d(b), M._d(c), len(t1),
                                      M._i32p(t1), M._i64p(l1), len(t2), M._i32p(t2), M._i64p(l2)))
        H.set_initial_iterate()
        ms, nbytes = H.time_matvec(2, 20)
        out["hybrid" if hyb else "dense"] = {"ms": ms, "bytes": nbytes, "kind": H.info("storage_kind"),
                                             "dense_equivalent_gbs": 8.0 * m * n / ms / 1e6}
    print(json.dumps({"check": "timing", "m": m, "n": n, **out}))


if __name__ == "__main__":
    t0 = time.time()
    cases = {"socls": problems.soc_constrained_ls(2100, 40, seed=3, scale=0.02),
             "socls_wide": problems.soc_constrained_ls(300, 200, seed=5, scale=0.02),
             "nnls": problems.nnls_conic(40, 50, seed=1, scale=0.02)}
    for label, P in cases.items():
        kind = operators(P, label)
        for alg in ("DR", "GAPA", "FISTA", "Dykstra"):
            lockstep(P, alg, label)
    timing(20000, 4000)
    print("hybrid_check done in %.1f s" % (time.time() - t0))
