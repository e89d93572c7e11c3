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


# ---------------------------------------------------------------------------------------------
# direct = true, LineSearchWrapper, the remaining cones (pins the oracle's restatements)
# ---------------------------------------------------------------------------------------------
def _dense_Q(P):
    A = np.asarray(P.A.todense()) if sp.issparse(P.A) else np.asarray(P.A)
    m, n = A.shape
    l = m + n + 1
    Q = np.zeros((l, l))
    Q[:n, n:n + m] = A.T
    Q[:n, -1] = P.c
    Q[n:n + m, :n] = -A
    Q[n:n + m, -1] = P.b
    Q[-1, :n] = -P.c
    Q[-1, n:n + m] = -P.b
    return Q


@pytest.mark.parametrize("kind", ["nnls", "sdp"])
def test_direct_projection_is_indaffine(kind):
    """HSDE.jl:10-15 / test/HSDEAffine.jl:72-80: with direct = true S1 is IndAffine([Q -I], 0); its prox is
    z - B'(B B')^-1 B z, lands on {Qu = v} and is idempotent."""
    P = problems.nnls_conic(12, 15, seed=1) if kind == "nnls" else problems.sdp_nearest_correlation(5, seed=4)
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=True)
    Q = _dense_Q(P)
    l = Q.shape[0]
    B = np.hstack([Q, -np.eye(l)])
    rng = np.random.default_rng(0)
    for _ in range(3):
        z = rng.standard_normal(2 * l)
        y = O.affine_prox(z)
        ref = z - B.T @ np.linalg.solve(B @ B.T, B @ z)
        assert np.abs(y - ref).max() < 1e-12 * max(1.0, np.abs(z).max())
        assert np.abs(B @ y).max() < 1e-12 * np.abs(z).max() * l
        assert np.abs(O.affine_prox(y) - y).max() < 1e-12


def test_direct_and_indirect_solves_agree_and_gapp_default_works():
    """test/testDRandGAPA.jl:37, test/testprint.jl:49: solvers constructed with direct=true (GAPP's default)
    reach the same optimum as the CG-based S1 and as SciPy's NNLS."""
    from scipy.optimize import nnls
    P = problems.nnls_conic(20, 25, seed=3)
    rng = np.random.default_rng(3)
    D, d = rng.standard_normal((20, 25)), rng.standard_normal(20)
    _, rn = nnls(D, d)
    sols = []
    for direct, alg in ((True, ("GAPP", 0.8, 1.8, 1.8, 0.0, 100)), (True, ("GAP", 0.5, 2.0, 2.0, 0.0, 100)),
                        (False, ("GAP", 0.5, 2.0, 2.0, 0.0, 100))):
        O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=direct)
        O.set_algorithm(*alg)
        O.set_iterate(O.initial_value())
        r = O.solve(max_iters=20000, checki=100, eps=1e-7)
        assert r["status"] == "Optimal", (direct, alg[0])
        x, y, s = O.populate_solution(r["guess"])
        assert abs(x[0] - rn) < 1e-4 * max(rn, 1.0)
        sols.append(x)
    assert np.abs(sols[1] - sols[2]).max() < 1e-4


def test_linesearch_wrapper_feasibility():
    """test/testfeasibility.jl:33-42 and test/linesearch.jl: LineSearchWrapper(GAP(eps=1e-8)) -> :Optimal,
    x >= 0, A x = b; on non-search iterations the wrapper is the inner algorithm."""
    A, b, cones = problems.feasibility_problem(50, 100, seed=2)
    O = fo.OracleFeasibility(A, b, np.zeros(100), 1, cones)
    O.set_algorithm("GAP", 0.8, 1.8, 1.8, 0.0, 100)
    O.set_linesearch(100)
    O.set_iterate(O.initial_value())
    r = O.solve(max_iters=10000, checki=100, eps=1e-8)
    assert r["status"] == "Optimal"
    x = r["guess"][:100]
    assert x.min() > -1e-12 and np.abs(A @ x - b).max() < 1e-6
    # lsinterval larger than the run: identical to the plain algorithm
    O1 = fo.OracleFeasibility(A, b, np.zeros(100), 1, cones)
    O2 = fo.OracleFeasibility(A, b, np.zeros(100), 1, cones)
    for O_, ls in ((O1, 0), (O2, 10 ** 6)):
        O_.set_algorithm("GAP", 0.8, 1.8, 1.8, 0.0, 100)
        O_.set_linesearch(ls)
        O_.set_iterate(O_.initial_value())
        O_.run(1, 50, checki=100, eps=1e-8)
    np.testing.assert_array_equal(O1.get_state("x"), O2.get_state("x"))


def test_rotated_soc_is_a_rotation_of_the_soc():
    """IndRotatedSOC (cones.jl:10; parity unpinned, published algorithm): equals R' P_SOC(R x) for the pi/4
    rotation R of the first two entries; the result satisfies 2 y1 y2 >= ||w||^2, y1, y2 >= 0."""
    rng = np.random.default_rng(0)
    c = np.sqrt(0.5)
    for _ in range(300):
        n = int(rng.integers(2, 9))
        x = rng.standard_normal(n) * rng.choice([0.1, 1, 10])
        y = fo.prox_cone("SOCRotated", x)
        R = np.eye(n)
        R[:2, :2] = [[c, c], [c, -c]]
        ref = R.T @ fo.prox_cone("SOC", R @ x)
        assert np.abs(y - ref).max() < 1e-13 * max(1.0, np.abs(x).max())
        assert y[0] >= -1e-12 and y[1] >= -1e-12
        assert 2 * y[0] * y[1] - np.sum(y[2:] ** 2) >= -1e-9 * max(1.0, np.abs(x).max()) ** 2
        yd = fo.prox_cone("SOCRotated", x, dual=True)                      # Moreau (cones.jl:80-85)
        assert np.abs(yd - (x + fo.prox_cone("SOCRotated", -x))).max() < 1e-15 * max(1.0, np.abs(x).max())


def test_exponential_cone_projection_properties():
    """IndExpPrimal / IndExpDual (cones.jl:12-13; parity unpinned, SCS's published projection restated):
    the result is in the cone, the map is idempotent to the algorithm's 1e-8 bisection accuracy, P_K(x) and
    P_K*(-x) are orthogonal and decompose x, and points already inside / in the polar cone are fixed / sent to 0."""
    rng = np.random.default_rng(0)

    def in_cone(v, tol):
        r, s, t = v
        return (s > 0 and s * np.exp(r / s) <= t + tol) or (r <= tol and abs(s) <= tol and t >= -tol)

    for _ in range(500):
        v = rng.standard_normal(3) * rng.choice([0.3, 1, 5])
        sc = max(1.0, np.abs(v).max())
        y = fo.prox_cone("ExpPrimal", v)
        assert in_cone(y, 1e-4 * sc)
        assert np.abs(fo.prox_cone("ExpPrimal", y) - y).max() < 1e-5 * sc
        pd = fo.prox_cone("ExpDual", -v)
        assert np.abs(y - pd - v).max() < 1e-12 * sc
        assert abs(np.dot(y, pd)) < 1e-5 * sc * sc
    np.testing.assert_array_equal(fo.prox_cone("ExpPrimal", np.array([0.0, 1.0, 2.0])), [0.0, 1.0, 2.0])   # inside
    np.testing.assert_array_equal(fo.prox_cone("ExpPrimal", np.array([-1.0, -1.0, 3.0])), [-1.0, 0.0, 3.0])  # r, s < 0
    np.testing.assert_array_equal(fo.prox_cone("ExpPrimal", np.array([1.0, -1.0, -5.0])), [0.0, 0.0, 0.0])   # polar
    with pytest.raises(NotImplementedError):
        fo.prox_cone("ExpPrimal", np.ones(4))


def test_feasibility_with_box():
    """test/testfeasibility.jl:9-19: S2 = IndBox(0, Inf) -- here through the oracle's box override."""
    A, b, _ = problems.feasibility_problem(30, 60, seed=4)
    O = fo.OracleFeasibility(A, b, np.zeros(60), 1, [("Free", 60), ("Zero", 30)])
    O.set_box(0, 60, 0.0, np.inf)
    O.set_algorithm("GAP", 0.5, 2.0, 2.0, 0.0, 100)
    O.set_iterate(O.initial_value())
    r = O.solve(max_iters=5000, checki=10, eps=1e-8)
    assert r["status"] == "Optimal"
    x = r["guess"][:60]
    assert x.min() > -1e-12 and np.abs(A @ x - b).max() < 1e-6


def test_c_and_numpy_oracles_agree_on_direct_and_linesearch():
    """Two independent restatements (C: hand-written Cholesky / loops; NumPy: LAPACK solve / vector ops) of the
    8f rows agree: direct = true trajectories to 1e-9 (no truncated CG in the way), LineSearchWrapper picks the
    same step lengths and stays within the CPU pair's usual drift."""
    P = problems.nnls_conic(12, 15, seed=2)
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=True)
    M = npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=True)
    for name, args in (("GAP", ("GAP", 0.5, 2.0, 2.0, 0.0, 100)), ("GAPA", ("GAPA", 1.0, 0.0, 0.0, 0.0, 100)),
                       ("Dykstra", ("Dykstra", 0.0, 0.0, 0.0, 0.0, 100))):
        O.set_algorithm(*args)
        M.set_algorithm(*args)
        O.set_iterate(O.initial_value())
        M.x = O.initial_value().copy()
        ro = O.solve(max_iters=60, checki=20, eps=1e-12)
        rm = M.solve(max_iters=60, checki=20, eps=1e-12)
        assert np.abs(ro["guess"] - rm["guess"]).max() < 1e-9 * np.abs(ro["guess"]).max(), name
        np.testing.assert_allclose(ro["history"]["p"], [h["p"] for h in rm["history"]], rtol=1e-7)
    # LineSearchWrapper(GAP) with the CG-based S1, well-conditioned instance
    P = problems.nnls_conic(12, 15, seed=2, scale=0.05)
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    M = npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    args = ("GAP", 0.8, 1.8, 1.8, 0.0, 100)
    O.set_algorithm(*args)
    O.set_linesearch(5)
    M.set_algorithm(*args)
    M.lsinterval = 5
    O.set_iterate(O.initial_value())
    M.x = O.initial_value().copy()
    M.checki, M.eps = 100, 1e-12
    for i in range(1, 16):
        O.run(i, 1, checki=100, eps=1e-12)
        M.i = i
        M.step()
        assert np.abs(M.x - O.get_state("x")).max() < 1e-6 * max(1.0, np.abs(M.x).max()), i
        assert O.s1_calls == M.S1.i
    assert fo.lib().fosor_get_alpha12 is not None
    fo.lib().fosor_get_alphabest.restype = __import__("ctypes").c_double
    fo.lib().fosor_get_alphabest.argtypes = [__import__("ctypes").c_void_p]
    assert fo.lib().fosor_get_alphabest(O._h) == M.alphabest


# ---------------------------------------------------------------------------------------------
# :Infeasible / :Unbounded branches of checkstatus (HSDEStatus.jl:57-66) -- no reference test reaches them
# ---------------------------------------------------------------------------------------------
STATUS_ALGS = {"DR": ("GAP", 0.5, 2.0, 2.0, 0.0, 100), "GAP": ("GAP", 0.8, 1.8, 1.8, 0.0, 100),
               "GAPA": ("GAPA", 1.0, 0.0, 0.0, 0.0, 100), "FISTA": ("FISTA", 1.0, 0.0, 0.0, 0.0, 100),
               "Dykstra": ("Dykstra", 0.0, 0.0, 0.0, 0.0, 100)}


@pytest.mark.parametrize("alg", sorted(STATUS_ALGS))
def test_infeasible_lp_is_reported_with_a_farkas_certificate(alg):
    P = problems.infeasible_lp()
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    O.set_algorithm(*STATUS_ALGS[alg])
    O.set_iterate(O.initial_value())
    r = O.solve(max_iters=3000, checki=50, eps=1e-6)
    assert r["status"] == "Infeasible"
    M = npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    M.set_algorithm(*STATUS_ALGS[alg])
    rn = M.solve(max_iters=3000, checki=50, eps=1e-6)
    assert (rn["status"], rn["iterations"]) == (r["status"], r["iterations"])
    # the certificate behind the decision (HSDEStatus.jl:62): A'y ~ 0 with b'y < 0, y in the dual cone
    l = P.m + P.n + 1
    y = r["guess"][P.n:P.n + P.m]
    assert y.min() >= 0 and P.b @ y < 0
    assert np.linalg.norm(P.A.T @ y) <= 1e-6 * (-(P.b @ y) / np.linalg.norm(P.b)) * 1.0001
    assert r["guess"][l - 1] < 1e-6          # tau -> 0


@pytest.mark.parametrize("alg", sorted(STATUS_ALGS))
def test_unbounded_lp_status_matches_between_restatements(alg):
    """GAP, GAPA, FISTA and Dykstra report :Unbounded with a recession direction.  DR(0.5) reports :Infeasible at
    its first check: the projected y is EXACTLY zero there, so the reference's test `norm(A'y) <= eps*(-b'y/norm(b))`
    reads `0.0 <= -0.0`, which is true (HSDEStatus.jl:62 evaluated literally).  Both restatements reproduce that."""
    P = problems.unbounded_lp()
    O = fo.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    O.set_algorithm(*STATUS_ALGS[alg])
    O.set_iterate(O.initial_value())
    r = O.solve(max_iters=3000, checki=50, eps=1e-6)
    M = npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    M.set_algorithm(*STATUS_ALGS[alg])
    rn = M.solve(max_iters=3000, checki=50, eps=1e-6)
    assert (rn["status"], rn["iterations"]) == (r["status"], r["iterations"])
    if alg == "DR":
        assert r["status"] == "Infeasible" and r["iterations"] == 50
        assert r["history"]["bty"][-1] == 0.0 and r["history"]["tau"][-1] == 0.0
    else:
        assert r["status"] == "Unbounded"
        x = r["guess"][:P.n]
        s = r["guess"][P.m + P.n + 1 + P.n:P.m + P.n + 1 + P.n + P.m]
        assert P.c @ x < 0                                                       # improving direction
        assert np.linalg.norm(P.A @ x + s) <= 1e-6 * (-(P.c @ x) / np.linalg.norm(P.c)) * 1.0001   # HSDEStatus.jl:60


def test_cg_iteration_cap_and_warning():
    """affinepluslinear.jl:115-120: max_iters = 1000; reaching it raises the @warn and S.cgiter == 1000."""
    A, b, cones, z = problems.stiff_feasibility_problem()
    O = fo.OracleFeasibility(A, b, np.zeros(A.shape[1]), 1, cones)
    y = O.affine_prox(z)
    assert O.cgiter == 1000 and fo.lib().fosor_get_cg_warned(O._h) == 1
    M = npo.NPModel.feasibility(A, b, np.zeros(A.shape[1]), 1, cones)
    yn = M.S1.prox(z)
    assert M.S1.cgiter == 1000
    assert np.abs(yn - y).max() < 1e-3 * np.abs(y).max()      # 1000 unconverged iterations: restatements drift to ~4e-6
    # a well-scaled instance stays far below the cap and leaves the flag alone
    A2, b2, cones2, z2 = problems.stiff_feasibility_problem(decades=0)
    O2 = fo.OracleFeasibility(A2, b2, np.zeros(A2.shape[1]), 1, cones2)
    O2.affine_prox(z2)
    assert O2.cgiter < 100 and fo.lib().fosor_get_cg_warned(O2._h) == 0
