"""The reference's own unit/integration tests, re-expressed against the CPU
oracle.  These pin the oracle: every property the reference asserts with a dense solve is
asserted here with NumPy's dense solve, plus the literal PSD known-answer of test/testPSD.jl.
(Julia's RNG is unavailable, so seeds are NumPy's; the properties are seed independent.)"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fos_oracle as fo
from oracle import np_oracle as npo

import fos_b200  # noqa: F401  (registers the package)
from fos_b200 import problems


# ---------------------------------------------------------------------------------------------
# test/conjugateGradient.jl
# ---------------------------------------------------------------------------------------------
def test_conjugate_gradient_spd():
    rng = np.random.default_rng(2)
    n = 300
    A = rng.random((n, n))
    A = A.T @ A                      # :6-7
    b = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    x, it = fo.cg_csc(A, b, x0, max_iters=100)          # :21
    assert it == 100
    x, it = fo.cg_csc(A, b, x, max_iters=5000)          # :23
    n1 = np.linalg.norm(A @ x - b)
    assert n1 < 1e-5                                     # :26
    xcopy = x + 1e-5 * rng.standard_normal(n)            # :28
    n2 = np.linalg.norm(A @ xcopy - b)
    xcopy, _ = fo.cg_csc(A, b, xcopy, max_iters=100)     # :30
    n3 = np.linalg.norm(A @ xcopy - b)
    assert n3 < 10 * n2                                  # :33


def test_cg_always_does_one_iteration():
    # conjugategradients.jl:37-43: the loop body runs before the first test
    A = np.eye(5)
    b = np.ones(5)
    x, it = fo.cg_csc(A, b, 0.9 * b, tol=1.0, max_iters=10)  # initial residual already below tol
    assert it == 1
    np.testing.assert_allclose(x, b, rtol=1e-15)


# ---------------------------------------------------------------------------------------------
# test/HSDEAffine.jl
# ---------------------------------------------------------------------------------------------
def _dense_q(A, b, c):
    m, n = A.shape
    Q = np.zeros((m + n + 1, m + n + 1))
    Q[:n, n:n + m] = A.T
    Q[:n, -1] = c
    Q[n:n + m, :n] = -A
    Q[n:n + m, -1] = b
    Q[-1, :n] = -c
    Q[-1, n:n + m] = -b
    return Q


@pytest.mark.parametrize("shape,density", [((100, 200), None), ((300, 600), 0.01)])
def test_hsde_q_and_matrix(shape, density):
    rng = np.random.default_rng(1)
    m, n = shape
    if density is None:
        A = rng.standard_normal((m, n))                  # :84-87
    else:
        A = sp.random(m, n, density=density, random_state=rng, data_rvs=rng.standard_normal).toarray()  # :89-90
    b = rng.standard_normal(m)
    c = rng.standard_normal(n)
    Q1 = _dense_q(A, b, c)
    O = fo.OracleConic(c, A, b, [("Free", m)], [("Free", n)])
    rhs = rng.standard_normal(m + n + 1)
    rhs_copy = rhs.copy()
    y2 = O.q_mul(rhs)
    assert np.array_equal(rhs, rhs_copy)                 # :35 input not mutated
    np.testing.assert_allclose(y2, Q1 @ rhs, rtol=1e-10, atol=1e-10)      # :36
    np.testing.assert_allclose(O.q_mul(rhs, transpose=True), Q1.T @ rhs, rtol=1e-10, atol=1e-10)  # :38-43
    l = m + n + 1
    M1 = np.block([[np.eye(l), Q1.T], [Q1, -np.eye(l)]])
    v = rng.standard_normal(2 * l)
    np.testing.assert_allclose(O.kkt_mul(v), M1 @ v, rtol=1e-10, atol=1e-10)                    # :45-62
    # HSDEMatrix.prox! == dense M\b with v <- Q u   (:71-81)
    bb = rng.standard_normal(2 * l)
    y2 = O.hsdematrix_prox(bb)
    y3 = np.linalg.solve(M1, bb)
    y3[l:] = Q1 @ y3[:l]
    np.testing.assert_allclose(y2, y3, rtol=1e-7, atol=1e-7)
    # and == projection onto {[u;v]: Qu = v} (IndAffine([Q -I], 0))
    B = np.hstack([Q1, -np.eye(l)])
    y1 = bb - B.T @ np.linalg.solve(B @ B.T, B @ bb)
    np.testing.assert_allclose(y2, y1, rtol=1e-7, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# test/affinepluslinear.jl
# ---------------------------------------------------------------------------------------------
def test_kkt_matrix_and_affinepluslinear():
    rng = np.random.default_rng(10)
    A = rng.standard_normal((10, 20))
    M1 = np.block([[np.eye(20), A.T], [A, -np.eye(10)]])
    x = rng.standard_normal(30)
    x0, z0 = rng.standard_normal(20), rng.standard_normal(10)
    q, b = rng.standard_normal(20), rng.standard_normal(10)
    cones = [("Free", 30)]
    O = fo.OracleFeasibility(A, b, q, 1, cones)
    np.testing.assert_allclose(O.kkt_mul(x), M1 @ x, rtol=1e-12, atol=1e-12)               # :7-19
    y2 = O.affine_prox(np.concatenate([x0, z0]))                                            # :28-47
    y3 = np.linalg.solve(M1, np.concatenate([x0 - q + A.T @ z0, b]))
    np.testing.assert_allclose(y2, y3, rtol=1e-9, atol=1e-9)
    O = fo.OracleFeasibility(A, b, q, -1, cones)                                            # :50-68
    y2 = O.affine_prox(np.concatenate([x0, z0]))
    M2 = np.block([[np.eye(20), -A.T], [A, np.eye(10)]])
    y3 = np.linalg.solve(M2, np.concatenate([x0 - q - A.T @ z0, b]))
    np.testing.assert_allclose(y2, y3, rtol=1e-9, atol=1e-9)
    # the prox is the projection: A x - beta z = b holds
    np.testing.assert_allclose(A @ y2[:20] + y2[20:], b, atol=1e-9)


# ---------------------------------------------------------------------------------------------
# test/testPSD.jl  -- the one literal known-answer in the reference
# ---------------------------------------------------------------------------------------------
YS = np.array([[-0.0064709, -0.22443], [-0.22443, -1.02411]])          # testPSD.jl:3-4
P_PSD_YS = np.array([[0.03909044662082823, -0.00823811392936668],
                     [-0.00823811392936668, 0.00173614084718757]])


def test_psd_literal_known_answer():
    w, V = np.linalg.eigh(YS)
    np.testing.assert_allclose(w, [-1.071407487468016, 0.0408265874680158], rtol=1e-12)
    v = problems.svec(YS)
    for impl in (lambda z: fo.prox_cone("SDP", z), npo.prox_sdp):
        P = problems.smat(impl(v))
        np.testing.assert_allclose(P, P_PSD_YS, rtol=1e-10, atol=1e-14)
    # dual through Moreau (cones.jl:80-85): PSD is self-dual
    Pd = problems.smat(fo.prox_cone("SDP", v, dual=True))
    np.testing.assert_allclose(Pd, P_PSD_YS, rtol=1e-10, atol=1e-14)


def test_psd_dr_solution_matches_projection():
    # testPSD.jl:22-25: minimize norm(vec(y - ys)) s.t. y PSD with DR(eps=1e-8) == projection (1e-8)
    P = problems.psd_projection_problem(YS)
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    O.set_algorithm("GAP", 0.5, 2.0, 2.0)
    O.set_iterate(O.initial_value())
    r = O.solve(max_iters=10000, checki=100, eps=1e-8)
    assert r["status"] == "Optimal"
    x, _, _ = O.populate_solution(r["guess"])
    np.testing.assert_allclose(problems.smat(x[1:]), P_PSD_YS, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# cones (ProximalOperators semantics, a17)
# ---------------------------------------------------------------------------------------------
def test_cone_projections_are_projections():
    rng = np.random.default_rng(0)
    for name in ("Free", "Zero", "NonNeg", "NonPos", "SOC", "SDP"):
        ln = 15 if name == "SDP" else 12
        for _ in range(5):
            x = rng.standard_normal(ln) * 3
            p = fo.prox_cone(name, x)
            np.testing.assert_allclose(fo.prox_cone(name, p), p, atol=1e-12)          # idempotent
            d = fo.prox_cone(name, x, dual=True)
            # Moreau: x = P_K(x) - P_K*(-x)  <=>  P_K*(x) = x + P_K(-x)
            np.testing.assert_allclose(d, x + fo.prox_cone(name, -x), atol=1e-12)
            np.testing.assert_allclose(p, npo.prox_cone(name, x), atol=1e-11)
            np.testing.assert_allclose(d, npo.prox_cone_dual(name, x), atol=1e-11)
            if name in ("SOC", "SDP", "NonNeg"):                                    # self-dual cones
                np.testing.assert_allclose(d, p, atol=1e-11)
                np.testing.assert_allclose(p @ (x - p), 0.0, atol=1e-10)              # orthogonality


def test_soc_cases():
    np.testing.assert_array_equal(fo.prox_cone("SOC", np.array([-5.0, 1, 2])), np.zeros(3))
    np.testing.assert_array_equal(fo.prox_cone("SOC", np.array([5.0, 1, 2])), [5.0, 1, 2])
    x = np.array([1.0, 3, 4])
    r = 0.5 * (1 + 1 / 5)
    np.testing.assert_allclose(fo.prox_cone("SOC", x), [r * 5, r * 3, r * 4], rtol=1e-15)


# ---------------------------------------------------------------------------------------------
# test/testfeasibility.jl (S1 = AffinePlusLinear in place of IndAffine, see problems.feasibility_problem)
# ---------------------------------------------------------------------------------------------
def _feas(alg_args, eps=1e-8, checki=100, max_iters=10000):
    A, b, cones = problems.feasibility_problem(50, 100, seed=2)
    O = fo.OracleFeasibility(A, b, np.zeros(100), 1, cones)
    O.set_algorithm(*alg_args)
    O.set_iterate(O.initial_value())
    r = O.solve(max_iters=max_iters, checki=checki, eps=eps)
    return A, b, r


def test_feasibility_dr_optimal():
    A, b, r = _feas(("GAP", 0.5, 2.0, 2.0), checki=10)                                   # :15
    assert r["status"] == "Optimal"
    x = r["guess"][:100]
    assert x.min() > -1e-12
    assert np.abs(A @ x - b).max() < 1e-6


@pytest.mark.parametrize("alg", [("GAPA", 1.0, 0.0, 0.0, 0.0), ("GAPP", 0.8, 1.8, 1.8, 0.0, 100)])
def test_feasibility_adaptive_solvers_optimal(alg):
    A, b, r = _feas(alg)                                                                   # :33-44
    assert r["status"] == "Optimal"
    x = r["guess"][:100]
    assert x.min() > -1e-12
    assert np.abs(A @ x - b).max() < 1e-6


def test_c_and_numpy_oracles_agree():
    P = problems.nnls_conic(12, 15, seed=7)
    for alg in (("GAP", 0.5, 2.0, 2.0), ("GAPA", 1.0, 0.0, 0.0, 0.0), ("FISTA", 1.0), ("Dykstra",),
                ("GAPP", 0.8, 1.8, 1.8, 0.0, 5)):
        O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
        O.set_algorithm(*alg)
        O.set_iterate(O.initial_value())
        r1 = O.run(1, 12, checki=4, eps=1e-9, trace=True)
        M = npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
        M.set_algorithm(*alg)
        r2 = M.solve(max_iters=12, checki=4, eps=1e-9, trace=True)
        # the truncated indefinite CG amplifies rounding ("parity budget"): 1e-7 over 12 steps
        np.testing.assert_allclose(r1["trace"], r2["trace"], rtol=0, atol=1e-7 * np.abs(r2["trace"]).max())
        np.testing.assert_allclose(r1["trace"][0], r2["trace"][0], rtol=0, atol=1e-10)
        assert list(r1["history"]["i"]) == [h["i"] for h in r2["history"]][:len(r1["history"]["i"])]


def test_hsde_status_and_solution_nnls():
    # README example shape (C1): GAP(0.5,2,2,max_iters=2000) on NNLS 40x50; optimum == scipy nnls
    from scipy.optimize import nnls
    P = problems.nnls_conic(40, 50, seed=1)
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    O.set_algorithm("GAP", 0.5, 2.0, 2.0)
    O.set_iterate(O.initial_value())
    r = O.solve(max_iters=2000, checki=100, eps=1e-5)
    assert r["status"] == "Optimal"
    x, y, s = O.populate_solution(r["guess"])
    rng = np.random.default_rng(1)
    D = rng.standard_normal((40, 50))
    d = rng.standard_normal(40)
    _, rn = nnls(D, d)
    assert abs(x[0] - rn) < 1e-4 * rn
    assert x[1:].min() > -1e-6
    # forced final check is absent when the last iteration was a check iteration (a-Q 3)
    assert r["history"]["i"][-1] == r["iterations"]
