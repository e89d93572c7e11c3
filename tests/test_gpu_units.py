"""GPU parity tests, unit level: every call goes through the C ABI (include/fos_b200.h) and is
compared with the CPU oracle on the same seeded inputs.  These mirror the reference's unit tests
(test/HSDEAffine.jl, test/affinepluslinear.jl, test/conjugateGradient.jl, test/testPSD.jl).

Tolerances (FP64): single operator applications 1e-12 relative (different summation order only);
a whole truncated CG solve 1e-9 (the indefinite KKT recurrence amplifies rounding ~1e4x,
DESIGN.md "parity budget")."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import load_affine, load_conic, rel_err

pytestmark = pytest.mark.gpu

OP_TOL = 1e-12
CG_TOL = 1e-9


def _rand_conic(problems, m, n, seed, density=None):
    h = m // 2
    return problems.random_feasible_conic(m, n, [("Zero", h), ("NonNeg", m - h)], seed=seed, density=density)


# shapes chosen to hit: tiny, ragged edges in both directions, exactly one tile, several bands,
# more row groups than CTAs
SHAPES = [(5, 3), (91, 51), (16, 512), (17, 513), (200, 2049), (1000, 300), (333, 4500), (2600, 2100)]


@pytest.mark.parametrize("m,n", SHAPES)
@pytest.mark.parametrize("path", ["tma", "plain", "sparse"])
def test_a_q_kkt_products(fos, oracle, m, n, path):
    from fos_b200 import problems
    P = _rand_conic(problems, m, n, seed=m * 7 + n)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    if path == "sparse":
        H = load_conic(fos, P, storage="sparse")
    else:
        H = load_conic(fos, P, storage="dense", matvec_impl=0 if path == "tma" else 1)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n)
    w = rng.standard_normal(m)
    assert rel_err(H.a_mul(x, m, n), O.a_mul(x)) < OP_TOL
    assert rel_err(H.a_mul(w, m, n, transpose=True), O.a_mul(w, transpose=True)) < OP_TOL
    B = rng.standard_normal(m + n + 1)
    assert rel_err(H.q_mul(B), O.q_mul(B)) < OP_TOL                      # HSDEAffine.jl:41-59
    assert rel_err(H.q_mul(B, transpose=True), O.q_mul(B, transpose=True)) < OP_TOL  # :61-65
    v = rng.standard_normal(2 * (m + n + 1))
    assert rel_err(H.kkt_mul(v), O.kkt_mul(v)) < OP_TOL                  # affinepluslinear.jl:37-49


def test_tma_and_plain_paths_agree_bitwise_shape(fos):
    """Both device paths on a multi-band, multi-CTA shape; the fused kernel must be deterministic."""
    from fos_b200 import problems
    P = _rand_conic(problems, 1500, 5000, seed=11)
    H1 = load_conic(fos, P, storage="dense", matvec_impl=0)
    H2 = load_conic(fos, P, storage="dense", matvec_impl=1)
    v = np.random.default_rng(5).standard_normal(2 * (P.m + P.n + 1))
    y1, y1b, y2 = H1.kkt_mul(v), H1.kkt_mul(v), H2.kkt_mul(v)
    assert np.array_equal(y1, y1b)          # fixed-order reductions: bitwise reproducible
    assert rel_err(y1, y2) < OP_TOL


def test_dense_load_direct_and_odd_lda(fos, oracle):
    from fos_b200 import problems
    P = _rand_conic(problems, 123, 777, seed=5)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    H = load_conic(fos, P, storage="dense_direct")
    v = np.random.default_rng(1).standard_normal(2 * (P.m + P.n + 1))
    assert rel_err(H.kkt_mul(v), O.kkt_mul(v)) < OP_TOL


@pytest.mark.parametrize("beta", [1, -1])
def test_affine_plus_linear_prox_matches_dense_solve(fos, oracle, beta):
    """test/affinepluslinear.jl:28-68"""
    rng = np.random.default_rng(10)
    A = rng.standard_normal((10, 20))
    x0, z0 = rng.standard_normal(20), rng.standard_normal(10)
    q, b = rng.standard_normal(20), rng.standard_normal(10)
    H = load_affine(fos, A, b, q, beta, [("Free", 30)])
    O = oracle.OracleFeasibility(A, b, q, beta, [("Free", 30)])
    xin = np.concatenate([x0, z0])
    H0 = load_affine(fos, A, b, q, beta, [("Free", 30)], fuse_rhs=0)   # reference-order rhs
    assert rel_err(H0.affine_prox(xin), H.affine_prox(xin)) < CG_TOL
    H = load_affine(fos, A, b, q, beta, [("Free", 30)])
    y = H.affine_prox(xin)
    if beta == 1:
        M = np.block([[np.eye(20), A.T], [A, -np.eye(10)]])
        y3 = np.linalg.solve(M, np.concatenate([x0 - q + A.T @ z0, b]))
    else:
        M = np.block([[np.eye(20), -A.T], [A, np.eye(10)]])
        y3 = np.linalg.solve(M, np.concatenate([x0 - q - A.T @ z0, b]))
    np.testing.assert_allclose(y, y3, rtol=1e-9, atol=1e-9)
    assert rel_err(y, O.affine_prox(xin)) < CG_TOL      # both converged to tol = an*eps
    assert abs(H.info("cgiter") - O.cgiter) <= 2         # the last iterations hover at the rounding floor
    assert H.info("s1_calls") == O.s1_calls == 2
    v = rng.standard_normal(30)
    assert rel_err(H.kkt_mul(v), O.kkt_mul(v)) < OP_TOL


@pytest.mark.parametrize("m,n,path", [(60, 90, "dense"), (60, 90, "sparse"), (300, 2100, "dense")])
def test_hsde_affine_prox_sequence(fos, oracle, m, n, path):
    """Three consecutive S1 proxes (warm start, decreasing tolerance 0.2^sqrt(i), call counter), on a
    well-conditioned instance (A scaled by 0.1) so that the truncated CG does not amplify rounding."""
    from fos_b200 import problems
    h = m // 2
    P = problems.random_feasible_conic(m, n, [("Zero", h), ("NonNeg", m - h)], seed=21, scale=0.1)
    for fuse in (0, 1):
        O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
        H = load_conic(fos, P, storage=path, fuse_rhs=fuse)
        rng = np.random.default_rng(2)
        for k in range(3):
            xin = rng.standard_normal(2 * (m + n + 1))
            yo = O.affine_prox(xin)
            # lock-step: same warm start and call counter on both sides
            if k > 0:
                H.set_info("s1_calls", O.s1_calls - 1)
            yg = H.affine_prox(xin)
            assert H.info("cgiter") == O.cgiter, f"CG iteration count differs at call {k}"
            assert rel_err(yg, yo) < CG_TOL
            if not fuse:  # the reference-order path materialises rhs (affinepluslinear.jl:94-95)
                assert rel_err(H.get_state("rhs"), O.get_state("rhs")) < OP_TOL
            H.set_state("xinit", O.get_state("xinit"))
        # fused: k+1 passes per projection, reference order: k+2
        assert H.info("total_passes") == H.info("total_cg") + (3 if fuse else 6)


def test_hsdematrix_prox_matches_dense_solve(fos, oracle):
    """test/HSDEAffine.jl:71-81: HSDEMatrix.prox! == dense M\\b with v <- Q u == projection onto {Qu = v}."""
    from fos_b200 import problems
    m, n = 40, 70
    P = problems.random_feasible_conic(m, n, [("Free", m)], seed=1)
    A = np.asarray(P.A)
    l = m + n + 1
    Q = np.zeros((l, l))
    Q[:n, n:n + m] = A.T
    Q[:n, -1] = P.c
    Q[n:n + m, :n] = -A
    Q[n:n + m, -1] = P.b
    Q[-1, :n] = -P.c
    Q[-1, n:n + m] = -P.b
    M1 = np.block([[np.eye(l), Q.T], [Q, -np.eye(l)]])
    H = load_conic(fos, P)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    bb = np.random.default_rng(3).standard_normal(2 * l)
    y = H.hsdematrix_prox(bb)
    y3 = np.linalg.solve(M1, bb)
    y3[l:] = Q @ y3[:l]
    np.testing.assert_allclose(y, y3, rtol=1e-8, atol=1e-8)
    B = np.hstack([Q, -np.eye(l)])
    y1 = bb - B.T @ np.linalg.solve(B @ B.T, B @ bb)          # IndAffine([Q -I], 0)
    np.testing.assert_allclose(y, y1, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(y, O.hsdematrix_prox(bb), rtol=1e-8, atol=1e-8)
    # S1's own state is untouched
    assert H.info("s1_calls") == 1 and H.info("cgiter") == 0
    v = np.random.default_rng(4).standard_normal(2 * l)
    assert rel_err(H.affine_prox(v), O.affine_prox(v)) < 1e-6


def test_cg_dense_spd(fos, oracle):
    """test/conjugateGradient.jl"""
    rng = np.random.default_rng(2)
    n = 300
    A = rng.random((n, n))
    A = A.T @ A
    b = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    H = fos.Handle(0)
    x, it = H.cg_dense(A, b, x0, max_iters=100)
    xo, ito = oracle.cg_csc(A, b, x0, max_iters=100)
    assert it == ito == 100
    x, it = H.cg_dense(A, b, x, max_iters=5000)
    n1 = np.linalg.norm(A @ x - b)
    assert n1 < 1e-5
    xcopy = x + 1e-5 * rng.standard_normal(n)
    n2 = np.linalg.norm(A @ xcopy - b)
    xcopy, _ = H.cg_dense(A, b, xcopy, max_iters=100)
    assert np.linalg.norm(A @ xcopy - b) < 10 * n2
    # short run, well conditioned: device and oracle agree closely
    A2 = np.eye(50) + 0.1 * (lambda G: G @ G.T)(rng.standard_normal((50, 50))) / 50
    b2 = rng.standard_normal(50)
    xg, itg = H.cg_dense(A2, b2, np.zeros(50), tol=1e-10, max_iters=200)
    xc, itc = oracle.cg_csc(A2, b2, np.zeros(50), tol=1e-10, max_iters=200)
    assert itg == itc
    assert rel_err(xg, xc) < 1e-12


# ---------------------------------------------------------------------------------------------
# cones
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,ln", [("Free", 37), ("Zero", 37), ("NonNeg", 1000), ("NonPos", 33), ("SOC", 1),
                                     ("SOC", 2), ("SOC", 3), ("SOC", 2050), ("SOC", 10001)])
@pytest.mark.parametrize("dual", [False, True])
def test_prox_cone_elementwise_and_soc(fos, oracle, name, ln, dual):
    H = fos.Handle(0)
    rng = np.random.default_rng(ln)
    for scale in (1.0, 10.0, 0.1):
        x = rng.standard_normal(ln)
        x[0] *= scale * (3 if name == "SOC" else 1)
        assert rel_err(H.prox_cone(name, x, dual), oracle.prox_cone(name, x, dual)) < 1e-13
    if name == "SOC" and ln > 1:
        for t in (-1e3, 1e3):  # the two trivial branches
            x = rng.standard_normal(ln)
            x[0] = t
            np.testing.assert_array_equal(H.prox_cone(name, x, dual), oracle.prox_cone(name, x, dual))


@pytest.mark.parametrize("ln", [2, 3, 4, 9, 2051, 6000])
@pytest.mark.parametrize("dual", [False, True])
def test_prox_rotated_soc(fos, oracle, ln, dual):
    """IndRotatedSOC (cones.jl:10): against the oracle, against R' P_SOC(R x) with the pi/4 rotation of the
    first two entries, and as a projection (membership 2 y1 y2 >= ||w||^2, idempotence)."""
    H = fos.Handle(0)
    rng = np.random.default_rng(ln)
    for scale in (1.0, 10.0, 0.1, -3.0):
        x = rng.standard_normal(ln)
        x[:2] *= scale
        y = H.prox_cone("SOCRotated", x, dual)
        assert rel_err(y, oracle.prox_cone("SOCRotated", x, dual)) < 1e-13
        if not dual:
            c = np.sqrt(0.5)
            Rx = x.copy()
            Rx[0], Rx[1] = c * x[0] + c * x[1], c * x[0] - c * x[1]
            ps = oracle.prox_cone("SOC", Rx)
            ref = ps.copy()
            ref[0], ref[1] = c * ps[0] + c * ps[1], c * ps[0] - c * ps[1]
            assert rel_err(y, ref) < 1e-12
            assert y[0] >= -1e-12 and y[1] >= -1e-12
            assert 2 * y[0] * y[1] - np.sum(y[2:] ** 2) >= -1e-9 * max(1.0, np.abs(x).max()) ** 2
            assert rel_err(H.prox_cone("SOCRotated", y), y) < 1e-12 or np.abs(y).max() < 1e-12


@pytest.mark.parametrize("name", ["ExpPrimal", "ExpDual"])
@pytest.mark.parametrize("dual", [False, True])
def test_prox_exponential_cones(fos, oracle, name, dual):
    """IndExpPrimal / IndExpDual (cones.jl:12-13), SCS's bisection + Newton projection: device == oracle (same
    algorithm; a bisection decision can flip on rounding, hence 1e-6), the result lies in the cone, the
    projection is idempotent to the algorithm's own accuracy and obeys Moreau's decomposition."""
    H = fos.Handle(0)
    rng = np.random.default_rng(7)
    X = rng.standard_normal((400, 3)) * rng.choice([0.3, 1.0, 5.0], size=(400, 1))
    X[:5] = [[1.0, 1.0, 3.0], [0.0, 0.0, 1.0], [-1.0, -1.0, 2.0], [-1.0, -2.0, -3.0], [2.0, 0.5, -1.0]]
    x = X.reshape(-1)                         # one cone entry holding 400 triples
    y = H.prox_cone(name, x, dual).reshape(-1, 3)
    yo = oracle.prox_cone(name, x, dual).reshape(-1, 3)
    assert np.abs(y - yo).max() <= 1e-6 * 5
    assert np.median(np.abs(y - yo)) < 1e-13
    if name == "ExpPrimal" and not dual:
        r, s, t = y[:, 0], y[:, 1], y[:, 2]
        with np.errstate(all="ignore"):
            inside = ((s > 0) & (s * np.exp(r / np.where(s > 0, s, 1.0)) <= t + 1e-4 * 5)) | \
                     ((r <= 1e-6) & (np.abs(s) <= 1e-6) & (t >= -1e-6))
        assert inside.all()
        y2 = H.prox_cone(name, y.reshape(-1)).reshape(-1, 3)
        assert np.abs(y2 - y).max() < 1e-5 * 5
        # Moreau: x = P_K(x) - P_K*(-x), with <P_K(x), P_K*(-x)> = 0
        pd = H.prox_cone("ExpDual", -x).reshape(-1, 3)
        assert np.abs(y - pd - X).max() < 1e-12 * 5
        assert np.abs(np.sum(y * pd, axis=1)).max() < 1e-5 * 25


def test_dual_cone_product_with_every_cone_type(fos, oracle):
    """cones.jl:122-142 with all nine cone types of conemap in K1 (supportedcones, FOSSolverInterface.jl:69)."""
    from fos_b200 import problems
    rng = np.random.default_rng(11)
    c1 = [("Zero", 4), ("SOCRotated", 7), ("ExpPrimal", 3), ("NonNeg", 5), ("ExpDual", 3), ("SOC", 6), ("SDP", 6),
          ("SOCRotated", 2), ("ExpPrimal", 3), ("NonPos", 2), ("Free", 3)]
    m = sum(l for _, l in c1)
    c2 = [("Free", 9), ("NonNeg", 4)]
    n = sum(l for _, l in c2)
    A = sp.random(m, n, density=0.4, random_state=rng, data_rvs=rng.standard_normal).tocsc()
    P = problems.ConicProblem(rng.standard_normal(n), A, rng.standard_normal(m), c1, c2)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    H = load_conic(fos, P)
    for _ in range(4):
        z = rng.standard_normal(2 * (m + n + 1)) * 2
        assert np.abs(H.cone_prox(z) - O.cone_prox(z)).max() < 1e-6      # exp cones: 1e-6, everything else 1e-12
    # a few GAP iterations run end to end in lock-step
    O.set_algorithm("GAP", 0.8, 1.8, 1.8, 0.0, 100)
    H.set_algorithm(fos.GAP())
    O.set_iterate(O.initial_value())
    H.ck(H.L.fos_begin_solve(H.h))
    for i in range(1, 6):
        H.set_state("x", O.get_state("x"))
        if O.s1_calls > 1:
            H.set_state("xinit", O.get_state("xinit"))
        H.set_info("s1_calls", O.s1_calls)
        O.run(i, 1, checki=100, eps=1e-9)
        H.run(i, 1, 100, 1e-9)
        assert rel_err(H.get_iterate(), O.get_state("x")) < 1e-5


def test_feasibility_with_indbox(fos, oracle):
    """test/testfeasibility.jl:9-19 shape: S2 = IndBox(0, Inf) on x (here also a two-sided box), S1 affine."""
    from fos_b200 import problems
    A, b, _ = problems.feasibility_problem(30, 60, seed=4)
    for lo, hi in ((0.0, np.inf), (0.0, 2.5)):
        cones = [("Box", 60, lo, hi), ("Zero", 30)]
        prob = fos.Feasibility(fos.AffinePlusLinear(A, b, np.zeros(60), 1), fos.ConeProduct(cones), 90)
        sol, model = fos.solve(prob, fos.DR(eps=1e-8, verbose=0), checki=10, max_iters=4000)
        O = oracle.OracleFeasibility(A, b, np.zeros(60), 1, [("Free", 60), ("Zero", 30)])
        O.set_box(0, 60, lo, hi)
        O.set_algorithm("GAP", 0.5, 2.0, 2.0, 0.0, 100)
        O.set_iterate(O.initial_value())
        ro = O.solve(max_iters=4000, checki=10, eps=1e-8)
        assert sol.status == ro["status"]
        assert abs(model.last_iteration - ro["iterations"]) <= 10
        if sol.status == "Optimal":
            x = sol.x[:60]
            assert x.min() > lo - 1e-9 and x.max() < hi + 1e-9
            assert np.abs(A @ x - b).max() < 1e-6


YS = np.array([[-0.0064709, -0.22443], [-0.22443, -1.02411]])          # test/testPSD.jl:3-4
P_PSD_YS = np.array([[0.03909044662082823, -0.00823811392936668],
                     [-0.00823811392936668, 0.00173614084718757]])


def test_psd_literal_known_answer(fos):
    from fos_b200 import problems
    H = fos.Handle(0)
    P = problems.smat(H.prox_cone("SDP", problems.svec(YS)))
    np.testing.assert_allclose(P, P_PSD_YS, rtol=1e-10, atol=1e-14)
    Pd = problems.smat(H.prox_cone("SDP", problems.svec(YS), dual=True))
    np.testing.assert_allclose(Pd, P_PSD_YS, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("d", [1, 2, 3, 5, 8, 15, 16, 17, 24, 33, 47, 48, 49, 64, 100, 112, 113, 130, 200, 256, 300, 512])
@pytest.mark.parametrize("dual", [False, True])
def test_psd_projection_vs_lapack(fos, d, dual):
    from oracle import np_oracle as npo
    H = fos.Handle(0)
    rng = np.random.default_rng(d)
    x = rng.standard_normal(d * (d + 1) // 2)
    ref = npo.prox_cone_dual("SDP", x) if dual else npo.prox_cone("SDP", x)
    got = H.prox_cone("SDP", x, dual)
    assert rel_err(got, ref) < 1e-12
    # degenerate spectrum: eigenvalues +-1 (a case one-sided Jacobi would get wrong)
    if d >= 2 and not dual:
        from fos_b200 import problems
        G, _ = np.linalg.qr(rng.standard_normal((d, d)))
        lam = np.where(np.arange(d) % 2 == 0, 1.0, -1.0)
        S = (G * lam) @ G.T
        got = problems.smat(H.prox_cone("SDP", problems.svec(S)))
        np.testing.assert_allclose(got, (G * np.maximum(lam, 0)) @ G.T, atol=1e-12)


def test_psd_mixed_orders_in_one_cone_set(fos):
    """One model whose constraint cones are SDP blocks of orders 3, 20, 7, 60, 16, 48, 112, 2 (in this order): the
    projection sorts them by order and hands them to the 32-, 128- and 512-thread kernels; every block must come back
    in its own place, primal and dual image alike (DualConeProduct, cones.jl:122-142)."""
    import scipy.sparse as sp
    from oracle import np_oracle as npo
    from helpers import load_conic
    from fos_b200.problems import ConicProblem
    ds = [3, 20, 7, 60, 16, 48, 112, 2]
    lens = [d * (d + 1) // 2 for d in ds]
    n = sum(lens)
    P = ConicProblem(np.zeros(n), -sp.identity(n, format="csc"), np.zeros(n), [("SDP", ln) for ln in lens], [("Free", n)])
    H = load_conic(fos, P, storage="sparse")
    rng = np.random.default_rng(5)
    z = rng.standard_normal(2 * (2 * n + 1))
    y = H.cone_prox(z)
    l = 2 * n + 1
    off = 0
    for ln in lens:
        a, b = n + off, l + n + off                       # y block (dual cone K1*), s block (K1)
        assert rel_err(y[a:a + ln], npo.prox_cone_dual("SDP", z[a:a + ln])) < 1e-12
        assert rel_err(y[b:b + ln], npo.prox_cone("SDP", z[b:b + ln])) < 1e-12
        off += ln


@pytest.mark.parametrize("d,nc", [(512, 2), (129, 5), (640, 1), (1024, 1)])
def test_psd_large_batched_cones(fos, d, nc):
    """Several large cones in one cooperative launch (config 4: the primal and the dual SDP(512) cone of
    DualConeProduct are projected together); spectra: random, low rank + noise, all negative, zero."""
    from oracle import np_oracle as npo
    from fos_b200 import problems
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
    Y, ms, sweeps = H.time_psd(X, reps=1)
    for k in range(nc):
        ref = npo.prox_cone("SDP", X[k])
        den = max(np.abs(ref).max(), np.abs(X[k]).max(), 1e-300)
        assert np.abs(Y[k] - ref).max() / den < 1e-12, (k, np.abs(Y[k] - ref).max() / den)
    assert 1 <= sweeps <= 40
    # idempotence and Moreau decomposition x = P_K(x) - P_K(-x) (size-independent properties)
    Y2, _, _ = H.time_psd(Y, reps=1)
    assert np.abs(Y2 - Y).max() <= 1e-12 * max(np.abs(Y).max(), 1.0)
    Yn, _, _ = H.time_psd(-X, reps=1)
    assert np.abs(Y - Yn - X).max() <= 1e-12 * np.abs(X).max()


@pytest.mark.parametrize("d", [640, 1024])
def test_psd_large_block_exchange_is_reproducible(fos, d):
    """The CTAs of a large cone exchange column blocks without flags: a reader recognises a block version by a bit in
    the data (csrc/psd_large.cu, "block exchange").  Accepting a stale version would not crash, it would change the
    result from run to run (it did, with two version buffers and 64 CTAs): repeated cold projections of the same
    matrix -- random, then rank-deficient, where sweeps are few and the CTAs drift apart most -- must agree bit for
    bit, take the same number of sweeps, and match LAPACK."""
    from oracle import np_oracle as npo
    rng = np.random.default_rng(7 * d)
    X = rng.standard_normal((1, d * (d + 1) // 2))
    ref = npo.prox_cone("SDP", X[0])
    first = None
    for rep in range(4):
        H = fos.Handle(0)
        Y, _, sw = H.time_psd(X, reps=2)
        Y2, _, sw2 = H.time_psd(Y, reps=2)
        if first is None:
            first = (Y.copy(), sw, Y2.copy(), sw2)
            assert np.abs(Y[0] - ref).max() <= 1e-12 * np.abs(ref).max()
            assert np.abs(Y2 - Y).max() <= 1e-12 * np.abs(Y).max()
        else:
            assert sw == first[1] and sw2 == first[3], (rep, sw, sw2, first[1], first[3])
            assert np.array_equal(Y, first[0]) and np.array_equal(Y2, first[2]), rep


def test_dual_cone_product_prox(fos, oracle):
    """cones.jl:122-142 on a mixed product: Zero + NonNeg + SOC + SOC + SDP rows, Free + NonNeg vars."""
    from fos_b200 import problems
    rng = np.random.default_rng(4)
    c1 = [("Zero", 7), ("NonNeg", 20), ("SOC", 9), ("SOC", 3), ("SDP", 10), ("NonPos", 4), ("Free", 3)]
    m = sum(l for _, l in c1)
    c2 = [("Free", 11), ("NonNeg", 6)]
    n = sum(l for _, l in c2)
    A = sp.random(m, n, density=0.3, random_state=rng, data_rvs=rng.standard_normal).tocsc()
    P = problems.ConicProblem(rng.standard_normal(n), A, rng.standard_normal(m), c1, c2)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    H = load_conic(fos, P)
    for _ in range(4):
        z = rng.standard_normal(2 * (m + n + 1)) * 2
        assert rel_err(H.cone_prox(z), O.cone_prox(z)) < 1e-12


def test_errors_are_loud(fos):
    from fos_b200 import problems
    H = fos.Handle(0)
    with pytest.raises(fos.FosError):
        H.get_iterate()                      # nothing loaded
    with pytest.raises(fos.FosError):
        H.prox_cone("ExpPrimal", np.ones(4))  # exponential cones come in triples
    P = problems.nnls_conic(4, 5, 1)
    P.constr_cones = [("SOC", 5)]            # does not cover 1:m
    with pytest.raises((fos.FosError, AssertionError)):
        load_conic(fos, P)


def test_sm_balanced_work_ranges(fos, oracle):
    """"k1_balance": the work ranges of the fused mat-vec are bound to SMs and sized to their measured speed.  Whatever
    the split, every row group is processed exactly once: the products agree with the oracle and with the even split;
    two handles of the same shape share the calibration (bit-identical results), repeated calls are bitwise equal."""
    from fos_b200 import problems
    P = _rand_conic(problems, 2512, 4100, seed=13)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    He = load_conic(fos, P, storage="dense_direct", k1_balance=0)
    Hb = load_conic(fos, P, storage="dense_direct", k1_balance=2)
    Hb2 = load_conic(fos, P, storage="dense_direct", k1_balance=2)
    assert He.info("k1_balanced") == 0
    if Hb.info("k1_balanced") != 1:
        pytest.skip("SM ids are not 0..G-1 on this device: the library kept the even split")
    rng = np.random.default_rng(3)
    for _ in range(3):
        v = rng.standard_normal(2 * (P.m + P.n + 1))
        yb = Hb.kkt_mul(v)
        assert rel_err(yb, O.kkt_mul(v)) < OP_TOL
        assert rel_err(yb, He.kkt_mul(v)) < OP_TOL
        assert np.array_equal(yb, Hb.kkt_mul(v))
        assert np.array_equal(yb, Hb2.kkt_mul(v))
    print(f"per-SM time spread (max - min) / mean: {Hb.info('k1_spread_before'):.3f} -> {Hb.info('k1_spread_after'):.3f}")
