"""GPU parity tests of batch mode (config 5: B independent conic problems of one shape, one
persistent CTA per problem, fos_*_batch of include/fos_b200.h) against the CPU oracle.

Same two kinds of comparison as tests/test_gpu_solvers.py: LOCK-STEP at 1e-10 on well-conditioned
instances (every problem of the batch receives ITS oracle's state before each iteration) and
FREE-RUNNING (status, iteration count, check iterations, solution), plus batch-specific properties:
problems stop independently, results do not depend on the number of persistent CTAs or on the
position of a problem inside the batch, and a batch of one reproduces the single-problem path.
"""
import numpy as np
import pytest

from helpers import ALG_SETUPS, load_conic, rel_err

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-10


def _dense_batch(plist):
    A = np.stack([np.asarray(P.A.todense()) if hasattr(P.A, "todense") else np.asarray(P.A) for P in plist])
    b = np.stack([P.b for P in plist])
    c = np.stack([P.c for P in plist])
    return A, b, c


def _load_batch(fos, plist, **options):
    H = fos.Handle(0)
    for k, v in options.items():
        H.set_option(k, v)
    A, b, c = _dense_batch(plist)
    H.load_conic_batch(A, b, c, plist[0].constr_cones, plist[0].var_cones)
    return H


def _nnls_batch(problems, B, rows, cols, scale, seed0=5):
    return [problems.nnls_conic(rows, cols, seed=seed0 + j, scale=scale) for j in range(B)]


def _sync_batch_from_oracles(H, Os, alg):
    H.set_state_batch("x", np.stack([O.get_state("x") for O in Os]))
    if Os[0].s1_calls > 1:
        H.set_state_batch("xinit", np.stack([O.get_state("xinit") for O in Os]))
    H.set_info_batch("s1_calls", [O.s1_calls for O in Os])
    if alg.startswith("GAPA"):
        H.set_info_batch("alpha12", [O.alpha12 for O in Os])
    if alg == "FISTA":
        H.set_state_batch("fista_y", np.stack([O.get_state("fista_y") for O in Os]))
        H.set_info_batch("fista_t", [O.fista_t for O in Os])
    if alg == "Dykstra":
        H.set_state_batch("dykstra_p", np.stack([O.get_state("dykstra_p") for O in Os]))
        H.set_state_batch("dykstra_q", np.stack([O.get_state("dykstra_q") for O in Os]))


@pytest.mark.parametrize("alg", ["FISTA", "Dykstra", "DR", "GAP", "GAPA", "GAPA_b", "AP"])
@pytest.mark.parametrize("shape", [(40, 50), (17, 70)])
def test_batch_lockstep_strict_1e10(fos, oracle, alg, shape):
    """Every problem of the batch, every iteration, from its oracle's state: next iterate within 1e-10,
    same CG iteration count, same S1 call counter, same p/d/g/ctx/bty/kappa/tau record."""
    from fos_b200 import problems
    B = 5
    plist = _nnls_batch(problems, B, shape[0], shape[1], scale=0.02)
    Os = [oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones) for P in plist]
    H = _load_batch(fos, plist)
    oargs, fac = ALG_SETUPS[alg]
    for O in Os:
        O.set_algorithm(*oargs)
        O.set_iterate(O.initial_value())
    H.set_algorithm(fac(fos))
    H.ck(H.L.fos_begin_solve_batch(H.h))
    checki, eps, worst = 5, 1e-12, 0.0
    for i in range(1, 31):
        _sync_batch_from_oracles(H, Os, alg)
        ros = [O.run(i, 1, checki=checki, eps=eps) for O in Os]
        done, st, recs = H.run_batch(i, 1, checki, eps)
        assert list(done) == [1] * B
        assert list(H.info_batch("cgiter")) == [O.cgiter for O in Os], f"iteration {i}: CG counts differ"
        assert list(H.info_batch("s1_calls")) == [O.s1_calls for O in Os]
        X = H.get_iterate_batch()
        for j, O in enumerate(Os):
            e = rel_err(X[j], O.get_state("x"))
            worst = max(worst, e)
            assert e < STEP_TOL, f"iteration {i}, problem {j}: iterate differs by {e:.3e}"
        if alg.startswith("GAPA"):
            np.testing.assert_allclose(H.info_batch("alpha12"), [O.alpha12 for O in Os], atol=1e-9)
        for j in range(B):
            ho = ros[j]["history"]
            if i % checki == 0:
                assert len(recs[j]) == 1 and recs[j][0, 0] == i
                for col, key in ((1, "p"), (2, "d"), (3, "g"), (4, "ctx"), (5, "bty"), (6, "kappa"), (7, "tau")):
                    np.testing.assert_allclose(recs[j][0, col], ho[key][0], rtol=1e-9, atol=1e-12, equal_nan=True,
                                               err_msg=key)
                assert recs[j][0, 8] == ho["cgiter"][0] and recs[j][0, 9] == ho["status"][0]
            else:
                assert len(recs[j]) == 0
    print(f"batch {shape} {alg}: worst one-step relative deviation {worst:.2e}")


@pytest.mark.parametrize("alg", ["FISTA", "Dykstra", "DR"])
def test_batch_free_running_solve_matches_oracle(fos, oracle, alg):
    """solve! of every problem in ONE launch: same status, iteration count and check iterations as the
    oracle; problems stop independently of each other."""
    from fos_b200 import problems
    B = 7
    plist = _nnls_batch(problems, B, 40, 50, scale=0.05, seed0=11)
    oargs, fac = ALG_SETUPS[alg]
    H = _load_batch(fos, plist)
    H.set_algorithm(fac(fos))
    eps = 1e-6 if alg == "DR" else 1e-4   # FISTA / Dykstra converge sublinearly (they are the slow C5 algorithms)
    max_iters = 1500
    done, st, recs, guess = H.solve_batch(max_iters, 50, eps)
    iters = []
    for j, P in enumerate(plist):
        O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
        O.set_algorithm(*oargs)
        O.set_iterate(O.initial_value())
        ro = O.solve(max_iters=max_iters, checki=50, eps=eps)
        assert fos.model.STATUS_SYMBOLS[st[j]] == ro["status"], j
        assert done[j] == ro["iterations"], (j, done[j], ro["iterations"])
        assert list(recs[j][:, 0]) == list(ro["history"]["i"])
        assert list(recs[j][:, 9]) == list(ro["history"]["status"])
        # free-running: the truncated CG's stop test flips now and then (parity budget)
        cg_g, cg_o = np.array(recs[j][:, 8]), np.array(ro["history"]["cgiter"])
        assert np.mean(cg_g != cg_o) <= 0.2 and np.abs(cg_g - cg_o).max() <= 4
        for col, key in ((1, "p"), (2, "d"), (3, "g")):
            np.testing.assert_allclose(recs[j][:, col], ro["history"][key], rtol=1e-2, atol=1e-2 * eps)
        assert rel_err(guess[j], ro["guess"]) < 1e-4
        iters.append(done[j])
    assert H.info_batch("s1_calls").tolist() == [d + 2 for d in iters]
  # iterations + getsol, counter starts at 1
    print(f"{alg}: iterations per problem {iters}")


def test_batch_of_one_equals_single_problem_path(fos):
    """The two device paths (multi-kernel single-problem, persistent batch) run the same algorithm: a batch
    of one follows the single handle step for step (reductions are associated differently: 1e-11)."""
    from fos_b200 import problems
    P = problems.nnls_conic(40, 50, seed=3, scale=0.05)
    for alg in ("FISTA", "Dykstra", "GAPA"):
        fac = ALG_SETUPS[alg][1]
        H1 = load_conic(fos, P, storage="dense")
        H1.set_algorithm(fac(fos))
        HB = _load_batch(fos, [P])
        HB.set_algorithm(fac(fos))
        H1.ck(H1.L.fos_begin_solve(H1.h))
        HB.ck(HB.L.fos_begin_solve_batch(HB.h))
        d1, s1, r1, _ = H1.run(1, 60, 20, 1e-12)
        db, sb, rb = HB.run_batch(1, 60, 20, 1e-12)
        assert d1 == db[0] == 60 and s1 == sb[0]
        assert rel_err(HB.get_iterate_batch()[0], H1.get_iterate()) < 1e-9
        np.testing.assert_allclose(rb[0][:, 1:8], r1[:, 1:8], rtol=1e-8, atol=1e-14)
        assert list(rb[0][:, 8]) == list(r1[:, 8])
        assert HB.info_batch("total_cg")[0] == H1.info("total_cg")
        g1, rec1, st1 = H1.finish()
        gb, recb, stb = HB.finish_batch()
        assert rel_err(gb[0], g1) < 1e-9 and st1 == stb[0] and len(rec1) == len(recb[0])


def test_batch_results_independent_of_cta_count_and_position(fos):
    """Bitwise reproducibility: a problem's trajectory does not depend on which CTA solves it, how many
    persistent CTAs there are, or where it sits in the batch (dynamic work counter, ragged stopping)."""
    from fos_b200 import problems
    B = 40
    plist = _nnls_batch(problems, B, 24, 30, scale=0.3, seed0=100)
    outs = []
    for ctas, order in ((0, None), (3, None), (7, np.random.default_rng(0).permutation(B))):
        pl = plist if order is None else [plist[k] for k in order]
        H = _load_batch(fos, pl, batch_ctas=ctas)
        H.set_algorithm(fos.FISTA())
        done, st, recs, guess = H.solve_batch(300, 25, 1e-3)
        if order is not None:
            inv = np.argsort(order)
            done, st, guess, recs = done[inv], st[inv], guess[inv], [recs[k] for k in inv]
        outs.append((done, st, guess, recs))
    for o in outs[1:]:
        np.testing.assert_array_equal(o[0], outs[0][0])
        np.testing.assert_array_equal(o[1], outs[0][1])
        np.testing.assert_array_equal(o[2], outs[0][2])
        for ra, rb in zip(o[3], outs[0][3]):
            np.testing.assert_array_equal(ra, rb)
    assert len(set(outs[0][0].tolist())) > 1, "test instance should stop at different iterations"


def test_batch_config5_shape(fos):
    """C5 shape (NNLS 256 x 512 -> m = 769, n = 513, lda = 528: nine consumer warps, the last one owning
    eight column pairs) on a few problems: FISTA, Dykstra and DR follow the single-problem device path
    (itself parity-tested against the oracle), and the record of the forced final check is reproduced
    by a NumPy evaluation of p and d on the returned point."""
    from fos_b200 import problems
    B = 3
    plist = _nnls_batch(problems, B, 256, 512, scale=1.0 / np.sqrt(512), seed0=5)
    for alg, fac in (("FISTA", lambda f: f.FISTA()), ("Dykstra", lambda f: f.Dykstra()), ("DR", lambda f: f.DR())):
        H = _load_batch(fos, plist)
        H.set_algorithm(fac(fos))
        done, st, recs, guess = H.solve_batch(130, 100, 1e-9)
        assert list(done) == [130] * B
        for j, P in enumerate(plist):
            m, n = P.m, P.n
            l = m + n + 1
            g = guess[j]
            tau = g[l - 1]
            x, y, s = g[:n] / tau, g[n:n + m] / tau, g[l + n:l + n + m] / tau
            A = np.asarray(P.A.todense())
            assert list(recs[j][:, 0]) == [100, 130]
            r = recs[j][-1]   # forced final check, evaluated on `guess` (solverwrapper.jl:31-34)
            p_np = np.linalg.norm(A @ x + s - P.b) / (1 + np.linalg.norm(P.b))
            d_np = np.linalg.norm(A.T @ y + P.c - g[l:l + n] / tau) / (1 + np.linalg.norm(P.c))
            np.testing.assert_allclose([r[1], r[2]], [p_np, d_np], rtol=1e-8, atol=1e-13)
            H1 = load_conic(fos, P, storage="dense")
            H1.set_algorithm(fac(fos))
            d1, s1, r1, g1 = H1.solve(130, 100, 1e-9)
            assert d1 == 130 and s1 == st[j]
            # unscaled instance: the two device paths drift apart like two CPU restatements do
            assert rel_err(g, g1) < 1e-3, (alg, j, rel_err(g, g1))
            np.testing.assert_allclose(recs[j][:, 1:8], r1[:, 1:8], rtol=2e-2, atol=1e-6)


def test_batch_hybrid_storage_equals_dense_storage(fos):
    """Rows with few non-zeros travel as CSR/CSC instead of dense tiles ("batch_hybrid", default on): the
    trajectories must equal the all-dense layout up to the association of the column sums, on matrices with
    ragged patterns (different dense/sparse/empty rows per problem)."""
    from fos_b200 import problems
    rng = np.random.default_rng(8)
    B, plist = 6, []
    for j in range(B):
        P = problems.nnls_conic(24, 30, seed=40 + j, scale=0.1)
        A = np.asarray(P.A.todense())
        A[1 + rng.integers(0, 24, size=j)] = 0.0              # a few empty rows, a different set per problem
        A[1 + rng.integers(0, 24), rng.integers(0, 31, size=28)] = 0.0  # a dense row turned sparse
        P.A = A
        plist.append(P)
    res = []
    for hybrid in (1, 0):
        H = _load_batch(fos, plist, batch_hybrid=hybrid)
        H.set_algorithm(fos.DR(0.5))
        H.ck(H.L.fos_begin_solve_batch(H.h))
        done, st, recs = H.run_batch(1, 40, 10, 1e-12)
        res.append((H.get_iterate_batch(), recs, H.info("bytes_per_pass"), H.info_batch("total_cg")))
    assert rel_err(res[0][0], res[1][0]) < 1e-9
    for ra, rb in zip(res[0][1], res[1][1]):
        np.testing.assert_allclose(ra[:, 1:8], rb[:, 1:8], rtol=1e-8, atol=1e-13)
    np.testing.assert_array_equal(res[0][3], res[1][3])
    assert res[0][2] < 0.6 * res[1][2]                      # the hybrid layout streams far fewer bytes


def test_batch_api_end_to_end_and_errors(fos):
    from fos_b200 import problems
    B = 4
    plist = _nnls_batch(problems, B, 20, 25, scale=0.2, seed0=30)
    A, b, c = _dense_batch(plist)
    models = fos.solve_batch(fos.DR(max_iters=800, eps=1e-6, checki=50), c, A, b, plist[0].constr_cones,
                             plist[0].var_cones)
    assert len(models) == B
    for mod in models:
        assert mod.solve_stat in ("Optimal", "Indeterminate")
        assert "p" in mod.history and "cgiter" in mod.history
        assert mod.primal_sol.shape == (26,)
    H = _load_batch(fos, plist)
    assert H.batch_size() == B
    with pytest.raises(fos.FosError):
        H.get_iterate()                       # single-problem entry point on a batch handle
    with pytest.raises(fos.FosError):
        H.set_algorithm(fos.GAPP(direct=False))   # not offered in batch mode
    H1 = load_conic(fos, plist[0])
    with pytest.raises(fos.FosError):
        H1.run_batch(1, 1, 1, 1e-3)           # batch entry point on a single-problem handle
