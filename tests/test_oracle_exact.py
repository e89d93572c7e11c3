"""CPU companion of tests/test_gpu_exact.py: the harness of the "exact" yardstick, exercised without a GPU.

`exact` = oracle/libfos_oracle_hp.so (long-double reductions), `C` = the restatement proper (the reference's
arithmetic), and in place of the CUDA path the independent NumPy restatement (pairwise sums, like the GPU's
tree sums).  Pins three facts the GPU test builds on:
  * with long-double reductions the restatement is unchanged as an algorithm (same CG counts on
    well-conditioned instances, iterates within 1e-12);
  * on the BASELINE-shaped instances the reference's own arithmetic is NOT within 1e-10 of exact
    (so no second implementation can be asked to be within 1e-10 of the reference there);
  * an implementation with tree-shaped sums is closer to exact than the sequential sums are.
"""
import numpy as np
import pytest

from helpers import ALG_SETUPS, assert_no_worse_than_reference_arithmetic, rel_err, three_way


@pytest.fixture(scope="module")
def problems():
    from fos_b200 import problems
    return problems


def test_hp_variant_is_the_same_algorithm(oracle, problems):
    assert oracle.lib("hp").fosor_variant() == 1 and oracle.lib("").fosor_variant() == 0
    P = problems.nnls_conic(40, 50, seed=1, scale=0.02)      # well-conditioned: no rounding amplification
    outs = []
    for v in ("", "hp"):
        O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, variant=v)
        O.set_algorithm(*ALG_SETUPS["DR"][0])
        O.set_iterate(O.initial_value())
        r = O.run(1, 60, checki=10, eps=1e-12)
        outs.append((O.get_iterate(), r["history"]["cgiter"], r["history"]["p"]))
    assert rel_err(outs[1][0], outs[0][0]) < 1e-12
    assert list(outs[1][1]) == list(outs[0][1])
    np.testing.assert_allclose(outs[1][2], outs[0][2], rtol=1e-8)


def test_set_state_round_trip(oracle, problems):
    P = problems.nnls_conic(12, 9, seed=3)
    A = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    B = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    for O in (A, B):
        O.set_algorithm(*ALG_SETUPS["GAPA"][0])
    A.set_iterate(A.initial_value())
    A.run(1, 7, checki=100, eps=1e-12)
    B.set_state("x", A.get_state("x"))
    B.set_state("xinit", A.get_state("xinit"))
    B.set_scalar("s1_calls", A.s1_calls)
    B.set_scalar("alpha12", A.alpha12)
    A.run(8, 3, checki=100, eps=1e-12)
    B.run(8, 3, checki=100, eps=1e-12)
    np.testing.assert_array_equal(A.get_iterate(), B.get_iterate())
    assert A.s1_calls == B.s1_calls and A.alpha12 == B.alpha12


@pytest.mark.parametrize("kind,alg", [("nnls", "DR"), ("lasso", "DR"), ("socls", "GAPA"), ("nnls", "Dykstra")])
def test_reference_arithmetic_is_not_within_1e10_of_exact(oracle, problems, kind, alg):
    from oracle import np_oracle as npo
    P = {"nnls": lambda: problems.nnls_conic(40, 50, seed=1), "lasso": lambda: problems.lasso_like(120, 260, seed=2),
         "socls": lambda: problems.soc_constrained_ls(2100, 40, seed=3)}[kind]()
    M = npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    M.set_algorithm(*ALG_SETUPS[alg][0])
    M.checki, M.eps = 100000, 1e-12

    def step_numpy(O, i):
        M.x = O.get_state("x").copy()
        if O.s1_calls > 1:
            M.S1.xinit = O.get_state("xinit").copy()
        M.S1.i = O.s1_calls
        M.alpha12, M.t = O.alpha12, O.fista_t
        M.y = O.get_state("fista_y").copy()
        M.p = O.get_state("dykstra_p").copy()
        M.q = O.get_state("dykstra_q").copy()
        M.i = i
        M.step()
        return M.x.copy(), M.S1.cgiter

    assert_no_worse_than_reference_arithmetic(f"{kind}/{alg}", *three_way(step_numpy, P, oracle, alg, 40),
                                              who="NumPy")
