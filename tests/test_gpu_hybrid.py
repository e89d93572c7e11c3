"""Hybrid row storage of the single-problem path ("hybrid_rows" = 1, csrc/matop.cu MatOp::init_hybrid): the contiguous
block of dense rows goes through the fused TMA mat-vec, the remaining non-empty rows (the -I blocks of SOC / NonNeg
constraints on the variables) as CSR + CSC, merged by one fold.  Operators against the oracle and against the
all-dense layout, then every algorithm in lock-step with the oracle at 1e-10 (well-conditioned instances)."""
import numpy as np
import pytest

from helpers import load_conic, rel_err, set_alg_both, sync_state_from_oracle

pytestmark = pytest.mark.gpu


def cases(problems):
    return {"socls_wide": problems.soc_constrained_ls(300, 200, seed=5, scale=0.02),   # dense block + 200 rows of -I
            "nnls": problems.nnls_conic(40, 50, seed=1, scale=0.02),                   # C1 / C5 structure
            "socls_tall": problems.soc_constrained_ls(2100, 40, seed=3, scale=0.02)}   # < 2 % to gain: stays dense


@pytest.mark.parametrize("label", ["socls_wide", "nnls", "socls_tall"])
def test_hybrid_operators_match_oracle_and_dense_layout(fos, oracle, label):
    from fos_b200 import problems
    P = cases(problems)[label]
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    Hd = load_conic(fos, P, storage="dense")
    Hh = load_conic(fos, P, storage="dense", hybrid_rows=1)
    kind = Hh.info("storage_kind")
    if label == "socls_tall":
        assert kind == 1 and Hh.info("bytes_per_pass") == Hd.info("bytes_per_pass")   # the plan declined
    else:
        assert kind == 3 and Hh.info("hybrid_sparse_rows") > 0
        assert Hh.info("bytes_per_pass") < 0.7 * Hd.info("bytes_per_pass")
    rng = np.random.default_rng(0)
    N = 2 * (P.m + P.n + 1)
    for _ in range(3):
        z = rng.standard_normal(N)
        kh, kd, ko = Hh.kkt_mul(z), Hd.kkt_mul(z), O.kkt_mul(z)
        assert rel_err(kh, ko) < 1e-11 and rel_err(kh, kd) < 1e-11
        ph, pd_, po = Hh.affine_prox(z), Hd.affine_prox(z), O.affine_prox(z)   # ONE call each: the prox advances S1.i
        assert rel_err(ph, pd_) < 1e-10 and rel_err(ph, po) < 1e-10
        x, y = rng.standard_normal(P.n), rng.standard_normal(P.m)
        assert rel_err(Hh.a_mul(x, P.m, P.n), O.a_mul(x)) < 1e-12
        assert rel_err(Hh.a_mul(y, P.m, P.n, transpose=True), O.a_mul(y, transpose=True)) < 1e-12


@pytest.mark.parametrize("label", ["socls_wide", "nnls"])
@pytest.mark.parametrize("alg", ["DR", "GAPA", "FISTA", "Dykstra"])
def test_hybrid_lockstep_1e10(fos, oracle, label, alg):
    from fos_b200 import problems
    P = cases(problems)[label]
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones)
    H = load_conic(fos, P, storage="dense", hybrid_rows=1)
    assert H.info("storage_kind") == 3
    set_alg_both(fos, H, O, alg)
    O.set_iterate(O.initial_value())
    H.ck(H.L.fos_begin_solve(H.h))
    for i in range(1, 31):
        sync_state_from_oracle(H, O, alg)
        ro = O.run(i, 1, checki=5, eps=1e-12)
        done, st, rec, _ = H.run(i, 1, 5, 1e-12)
        assert H.info("cgiter") == O.cgiter, (label, alg, i)
        assert rel_err(H.get_iterate(), O.get_state("x")) < 1e-10, (label, alg, i)
        if i % 5 == 0:
            np.testing.assert_allclose(rec[0, 1:8], [ro["history"][k][0] for k in ("p", "d", "g", "ctx", "bty", "kappa", "tau")],
                                       rtol=1e-9, atol=1e-12, equal_nan=True)
