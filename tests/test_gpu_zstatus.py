"""GPU parity of the :Infeasible / :Unbounded branches of checkstatus (HSDEStatus.jl:57-66, k6_check_hsde)
against the CPU oracle, on LPs with a built-in Farkas certificate / recession direction.  No reference test
reaches these branches; the oracle's behaviour on them is pinned in tests/test_oracle_reference_tests.py
(C and NumPy restatements agree, certificates verified), including the literal `0.0 <= -0.0` outcome of
HSDEStatus.jl:62 that makes DR(0.5) report :Infeasible on the unbounded instance."""
import numpy as np
import pytest

from helpers import ALG_SETUPS, load_conic, rel_err, set_alg_both, sync_state_from_oracle

pytestmark = pytest.mark.gpu

CHECKI, EPS, MAX_ITERS = 50, 1e-6, 3000


def _problem(kind):
    from fos_b200 import problems
    return problems.infeasible_lp() if kind == "infeasible" else problems.unbounded_lp()


@pytest.mark.parametrize("kind", ["infeasible", "unbounded"])
@pytest.mark.parametrize("alg", ["DR", "GAP", "GAPA", "FISTA", "Dykstra"])
def test_status_branches_lockstep(fos, oracle, kind, alg):
    """From the oracle's state, every iteration up to the oracle's decision: same status code at every check
    (Continue, then Infeasible / Unbounded at the same iteration), kappa / tau and the objective terms agree."""
    P = _problem(kind)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    H = load_conic(fos, P)
    set_alg_both(fos, H, O, alg)
    O.set_iterate(O.initial_value())
    H.ck(H.L.fos_begin_solve(H.h))
    decided = None
    for i in range(1, MAX_ITERS + 1):
        sync_state_from_oracle(H, O, alg)
        ro = O.run(i, 1, checki=CHECKI, eps=EPS)
        done, st, rec, _ = H.run(i, 1, CHECKI, EPS)
        assert done == 1
        assert rel_err(H.get_iterate(), O.get_state("x")) < 1e-6, i    # unscaled instance: see DESIGN.md, parity budget
        if i % CHECKI == 0:
            ho = ro["history"]
            assert len(rec) == 1 and rec[0, 0] == i
            assert rec[0, 9] == ho["status"][0], f"iteration {i}: status {rec[0, 9]} vs {ho['status'][0]}"
            for col, key in ((4, "ctx"), (5, "bty"), (6, "kappa"), (7, "tau")):
                np.testing.assert_allclose(rec[0, col], ho[key][0], rtol=1e-5, atol=1e-9, err_msg=f"{key} at {i}")
            if ho["status"][0] != 0:
                decided = (i, int(ho["status"][0]))
                break
    assert decided is not None, "the oracle did not reach a decision"
    want = {"infeasible": 3, "unbounded": 3 if alg == "DR" else 2}[kind]   # DR: the 0.0 <= -0.0 case (module docstring)
    assert decided[1] == want


@pytest.mark.parametrize("kind", ["infeasible", "unbounded"])
@pytest.mark.parametrize("alg", ["DR", "GAPA"])
def test_status_branches_free_running(fos, oracle, kind, alg):
    """solve!(model) end to end: the same status symbol, found within one check interval of the oracle's."""
    P = _problem(kind)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    H = load_conic(fos, P)
    set_alg_both(fos, H, O, alg)
    O.set_iterate(O.initial_value())
    ro = O.solve(max_iters=MAX_ITERS, checki=CHECKI, eps=EPS)
    done, st, rec, guess = H.solve(MAX_ITERS, CHECKI, EPS)
    assert fos.model.STATUS_SYMBOLS[st] == ro["status"]
    assert abs(done - ro["iterations"]) <= CHECKI


def test_cg_iteration_cap_sets_the_warning(fos, oracle):
    """affinepluslinear.jl:115-120 / conjugategradients.jl:43,53: the solve stops at max_iters = 1000, S.cgiter is
    1000 and the @warn is raised (fos_get_info 4); a well-scaled instance leaves the flag alone."""
    from fos_b200 import problems
    from helpers import load_affine
    A, b, cones, z = problems.stiff_feasibility_problem()
    O = oracle.OracleFeasibility(A, b, np.zeros(A.shape[1]), 1, cones)
    yo = O.affine_prox(z)
    assert O.cgiter == 1000
    H = load_affine(fos, A, b, np.zeros(A.shape[1]), 1, cones)
    y = H.affine_prox(z)
    assert H.info("cgiter") == 1000
    assert H.info("cg_warned") == 1
    assert rel_err(y, yo) < 1e-2       # 1000 unconverged CG iterations: two CPU restatements drift to 4e-6
    A2, b2, cones2, z2 = problems.stiff_feasibility_problem(decades=0)
    H2 = load_affine(fos, A2, b2, np.zeros(A2.shape[1]), 1, cones2)
    O2 = oracle.OracleFeasibility(A2, b2, np.zeros(A2.shape[1]), 1, cones2)
    assert rel_err(H2.affine_prox(z2), O2.affine_prox(z2)) < 1e-10
    assert H2.info("cgiter") == O2.cgiter and H2.info("cg_warned") == 0


@pytest.mark.parametrize("alg", ["DR", "GAPA", "Dykstra"])
def test_status_branches_batch_mode(fos, oracle, alg):
    """Batch mode (one persistent CTA per problem, bt_check of csrc/batch.cu): a batch that mixes infeasible and
    unbounded LPs of one shape reports, per problem, the oracle's status within one check interval -- problems
    stop independently."""
    from fos_b200 import problems
    # m = 91, n = 51: the shape of config 1 / the batch lock-step tests (tests/test_gpu_batch.py)
    plist = [problems.infeasible_lp(91, 51, seed=7), problems.unbounded_lp(91, 51, seed=8),
             problems.infeasible_lp(91, 51, seed=9), problems.unbounded_lp(91, 51, seed=14)]
    A = np.stack([np.asarray(P.A.todense()) for P in plist])
    b = np.stack([P.b for P in plist])
    c = np.stack([P.c for P in plist])
    H = fos.Handle(0)
    H.load_conic_batch(A, b, c, plist[0].constr_cones, plist[0].var_cones)
    oargs, fac = ALG_SETUPS[alg]
    H.set_algorithm(fac(fos))
    done, st, recs, guess = H.solve_batch(MAX_ITERS, CHECKI, EPS)
    for j, P in enumerate(plist):
        O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
        O.set_algorithm(*oargs)
        O.set_iterate(O.initial_value())
        ro = O.solve(max_iters=MAX_ITERS, checki=CHECKI, eps=EPS)
        assert ro["status"] in ("Infeasible", "Unbounded"), (j, ro["status"])
        assert fos.model.STATUS_SYMBOLS[st[j]] == ro["status"], j
        assert abs(done[j] - ro["iterations"]) <= CHECKI, (j, done[j], ro["iterations"])
