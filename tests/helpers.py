"""Shared helpers of the parity tests: load the same (c, A, b, cones) into the CUDA library
(through the C ABI) and into the oracle."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

CONE = {"Free": 0, "Zero": 1, "NonNeg": 2, "NonPos": 3, "SOC": 4, "SOCRotated": 5, "SDP": 6}


def rel_err(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    den = max(np.abs(b).max(initial=0.0), 1e-300)
    return float(np.abs(a - b).max(initial=0.0) / den)


def _tool_overrides(options):
    """FOS_TEST_USE_GRAPHS=0 runs every test on the kernel-per-launch path unless the test chooses itself:
    compute-sanitizer's synccheck / racecheck cannot follow a conditional WHILE node (scripts/gpu_sanitize_r2.sh,
    profiles/r2_sanitizer.md), so those tools check the same kernels launched one by one."""
    import os
    v = os.environ.get("FOS_TEST_USE_GRAPHS")
    if v is not None and "use_graphs" not in options:
        options["use_graphs"] = int(v)


def load_conic(fos, P, storage="auto", **options):
    """ConicProblem -> fos Handle with the problem loaded (C ABI: fos_load_conic_csc/dense)."""
    from fos_b200 import model as M
    H = fos.Handle(0)
    _tool_overrides(options)
    for k, v in options.items():
        H.set_option(k, v)
    t1, l1 = M._cone_arrays(P.constr_cones, P.m, "constraint")
    t2, l2 = M._cone_arrays(P.var_cones, P.n, "variable")
    b = np.ascontiguousarray(P.b, float)
    c = np.ascontiguousarray(P.c, float)
    if storage == "dense_direct":
        Ad = np.ascontiguousarray(P.A.toarray() if sp.issparse(P.A) else P.A, dtype=np.float64)
        H.ck(H.L.fos_load_conic_dense(H.h, P.m, P.n, Ad.ctypes.data_as(C.c_void_p), P.n, 0, 0, P.m, M._d(b),
                                      M._d(c), len(t1), M._i32p(t1), M._i64p(l1), len(t2), M._i32p(t2),
                                      M._i64p(l2)))
    else:
        _, colptr, rowval, nzval = M._csc_arrays(P.A)
        code = {"auto": 0, "dense": 1, "sparse": 2}[storage]
        H.ck(H.L.fos_load_conic_csc(H.h, P.m, P.n, M._i64p(colptr), M._i64p(rowval), M._d(nzval), 0, M._d(b),
                                    M._d(c), len(t1), M._i32p(t1), M._i64p(l1), len(t2), M._i32p(t2), M._i64p(l2),
                                    code))
    H.set_initial_iterate()
    return H


def load_affine(fos, A, b, q, beta, cones, decreasing=False, storage="auto", **options):
    from fos_b200 import model as M
    H = fos.Handle(0)
    _tool_overrides(options)
    for k, v in options.items():
        H.set_option(k, v)
    Am, colptr, rowval, nzval = M._csc_arrays(A)
    am, an = Am.shape
    t, ln = M._cone_arrays(cones, am + an, "S2")
    b = np.ascontiguousarray(b, float)
    q = np.ascontiguousarray(q, float)
    code = {"auto": 0, "dense": 1, "sparse": 2}[storage]
    H.ck(H.L.fos_load_affine_csc(H.h, am, an, M._i64p(colptr), M._i64p(rowval), M._d(nzval), 0, M._d(b), M._d(q),
                                 int(beta), 1 if decreasing else 0, len(t), M._i32p(t), M._i64p(ln), code))
    H.set_initial_iterate()
    return H


ALG_SETUPS = {
    # name: (oracle args, fos algorithm factory)
    "DR": (("GAP", 0.5, 2.0, 2.0, 0.0, 100), lambda f: f.DR(0.5)),
    "GAP": (("GAP", 0.8, 1.8, 1.8, 0.0, 100), lambda f: f.GAP()),
    "AP": (("GAP", 1.0, 1.0, 1.0, 0.0, 100), lambda f: f.AP()),
    "GAPA": (("GAPA", 1.0, 0.0, 0.0, 0.0, 100), lambda f: f.GAPA()),
    "GAPA_b": (("GAPA", 0.8, 0.0, 0.0, 0.9, 100), lambda f: f.GAPA(0.8, 0.9)),
    "FISTA": (("FISTA", 1.0, 0.0, 0.0, 0.0, 100), lambda f: f.FISTA()),
    "Dykstra": (("Dykstra", 0.0, 0.0, 0.0, 0.0, 100), lambda f: f.Dykstra()),
    "GAPP": (("GAPP", 0.8, 1.8, 1.8, 0.0, 7), lambda f: f.GAPP(direct=False, iproj=7)),
}


def set_alg_both(fos, H, O, name):
    oargs, fac = ALG_SETUPS[name]
    O.set_algorithm(*oargs)
    H.set_algorithm(fac(fos))


def sync_state_from_oracle(H, O, alg):
    """Copy every persistent piece of solver state from the oracle into the GPU handle so that the
    next iteration starts from bit-identical inputs (lock-step parity)."""
    H.set_state("x", O.get_state("x"))
    if O.s1_calls > 1:
        H.set_state("xinit", O.get_state("xinit"))
    H.set_info("s1_calls", O.s1_calls)
    if alg.startswith("GAPA"):
        H.set_info("alpha12", O.alpha12)
    if alg == "FISTA":
        H.set_state("fista_y", O.get_state("fista_y"))
        H.set_info("fista_t", O.fista_t)
    if alg == "Dykstra":
        H.set_state("dykstra_p", O.get_state("dykstra_p"))
        H.set_state("dykstra_q", O.get_state("dykstra_q"))


# ---------------------------------------------------------------------------------------------
# the "exact" yardstick (tests/test_gpu_exact.py, tests/test_oracle_exact.py)
# ---------------------------------------------------------------------------------------------
def sync_oracle_from_oracle(dst, O):
    """Put the `exact` restatement into the C oracle's state."""
    dst.set_state("x", O.get_state("x"))
    if O.s1_calls > 1:
        dst.set_state("xinit", O.get_state("xinit"))
    dst.set_scalar("s1_calls", O.s1_calls)
    dst.set_scalar("alpha12", O.alpha12)
    dst.set_scalar("fista_t", O.fista_t)
    dst.set_state("fista_y", O.get_state("fista_y"))
    dst.set_state("dykstra_p", O.get_state("dykstra_p"))
    dst.set_state("dykstra_q", O.get_state("dykstra_q"))


def three_way(step_other, P, oracle, alg, n_iter):
    """Runs C / exact / `other` in lock-step; returns (e_c, e_other, flips_c, flips_other) against exact."""
    oargs = ALG_SETUPS[alg][0]
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    X = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, variant="hp")
    O.set_algorithm(*oargs)
    X.set_algorithm(*oargs)
    O.set_iterate(O.initial_value())
    e_c, e_o, f_c, f_o = [], [], 0, 0
    for i in range(1, n_iter + 1):
        sync_oracle_from_oracle(X, O)
        x_other, cg_other = step_other(O, i)
        O.run(i, 1, checki=100000, eps=1e-12)
        X.run(i, 1, checki=100000, eps=1e-12)
        xx = X.get_state("x")
        fc, fo_ = O.cgiter != X.cgiter, cg_other != X.cgiter
        f_c += fc
        f_o += fo_
        if not fc and not fo_:
            e_c.append(max(rel_err(O.get_state("x"), xx), 1e-17))
            e_o.append(max(rel_err(x_other, xx), 1e-17))
    return np.array(e_c), np.array(e_o), f_c, f_o


def assert_no_worse_than_reference_arithmetic(tag, e_c, e_o, f_c, f_o, min_iters=6, who="GPU", premise=True):
    ratio = e_o / e_c
    gm = float(np.exp(np.mean(np.log(ratio))))
    print(f"{tag}: C-vs-exact median {np.median(e_c):.2e} max {e_c.max():.2e} flips {f_c} | "
          f"{who}-vs-exact median {np.median(e_o):.2e} max {e_o.max():.2e} flips {f_o} | "
          f"ratio geo-mean {gm:.2f} max {ratio.max():.1f} over {len(ratio)} iterations")
    assert len(ratio) >= min_iters
    assert f_o <= f_c + 2, "the CUDA path misses the exact CG count more often than the reference's arithmetic"
    assert gm <= 1.0, "the CUDA path is farther from the exact iteration than the reference's own arithmetic"
    assert np.median(e_o) <= 2.0 * np.median(e_c)
    assert e_o.max() <= 50.0 * e_c.max()
    # and the premise of the test: the reference's arithmetic is itself NOT within 1e-10 of exact here
    if premise:
        assert e_c.max() > 1e-10


