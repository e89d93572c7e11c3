"""Tests against the committed fixtures of tests/golden/ (made by tests/golden/make_golden.py).

* reference_literals.json: literal values of the reference's own tests / sources.
* lockstep_*.npz: frozen trajectories of the CPU oracle (ORACLE outputs -- the Julia reference cannot run
  here).  CPU: the C oracle regenerates them, the independent NumPy restatement reproduces every step
  from the stored state.  GPU (through the C ABI): from the stored state before iteration i, one
  iteration reproduces the stored result to 1e-10 with the same CG count and check record.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import ALG_SETUPS, load_conic, rel_err

GOLDEN = Path(__file__).resolve().parent / "golden"
FIXTURES = sorted(p.name for p in GOLDEN.glob("lockstep_*.npz"))
STEP_TOL = 1e-10  # north_star: per-iteration iterates within 1e-10 relative


def _literals():
    return json.loads((GOLDEN / "reference_literals.json").read_text())


def _load(name):
    from fos_b200 import problems
    z = np.load(GOLDEN / name, allow_pickle=False)
    shape = tuple(int(v) for v in z["A_shape"])
    A = sp.csc_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=shape)
    cones = json.loads(str(z["cones"]))
    P = problems.ConicProblem(c=z["c"], A=A, b=z["b"], constr_cones=[(n, k) for n, k in cones["constr"]],
                              var_cones=[(n, k) for n, k in cones["var"]], name=name)
    return z, P, str(z["alg"])


def _direct(z):
    return bool(int(z["direct"])) if "direct" in z.files else False


def _tol(z, alg):
    # GAPP's projected step multiplies a difference of projections by alpha_best = 2^k (gapproj.jl:46-58); the
    # direct fixtures are UNSCALED instances (as tests/test_gpu_solvers.py::test_direct_lockstep_and_solve)
    if alg == "GAPP":
        return STEP_TOL * (100 if _direct(z) else 10)
    return STEP_TOL


def test_fixtures_are_present():
    assert (GOLDEN / "reference_literals.json").exists()
    assert len(FIXTURES) >= 10


# ---------------------------------------------------------------------------------------------
# literals of the reference
# ---------------------------------------------------------------------------------------------
def test_literal_psd_projection(oracle):
    """test/testPSD.jl:3-4,14-19: the projection of the literal 2x2 matrix."""
    from fos_b200 import problems
    from oracle import np_oracle as npo
    lit = _literals()["testPSD"]
    ys, want = np.array(lit["ys"]), np.array(lit["projection"])
    np.testing.assert_allclose(np.linalg.eigvalsh(ys), lit["eigenvalues"], rtol=1e-12)
    w, V = np.linalg.eigh(ys)
    np.testing.assert_allclose((V * np.maximum(w, 0)) @ V.T, want, rtol=1e-10, atol=1e-14)
    v = problems.svec(ys)
    np.testing.assert_allclose(problems.smat(oracle.prox_cone("SDP", v)), want, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(problems.smat(npo.prox_sdp(v)), want, rtol=1e-10, atol=1e-14)


def test_literal_constructor_defaults(fos):
    """solvers/gap.jl:13, solvers.jl:10-11, gapa.jl:16, fista.jl:10, dykstra.jl:9, gapproj.jl:14; solverwrapper.jl:5-9."""
    d = _literals()["defaults"]
    g = fos.GAP()
    assert (g.α, g.α1, g.α2, g.direct) == (d["GAP"]["alpha"], d["GAP"]["alpha1"], d["GAP"]["alpha2"], False)
    r = fos.DR()
    assert (r.α, r.α1, r.α2) == (d["DR"]["alpha"], d["DR"]["alpha1"], d["DR"]["alpha2"])
    a = fos.AP()
    assert (a.α, a.α1, a.α2) == (d["AP"]["alpha"], d["AP"]["alpha1"], d["AP"]["alpha2"])
    ga = fos.GAPA()
    assert (ga.α, ga.β, ga.direct) == (d["GAPA"]["alpha"], d["GAPA"]["beta"], False)
    assert fos.FISTA().α == d["FISTA"]["alpha"] and fos.FISTA().direct is False
    assert fos.Dykstra().direct is False
    gp = fos.GAPP()
    assert (gp.α, gp.α1, gp.α2, gp.iproj, gp.direct) == (0.8, 1.8, 1.8, d["GAPP"]["iproj"], True)


@pytest.mark.gpu
def test_literal_psd_projection_gpu(fos):
    """The same literal through the C ABI (fos_prox_cone, K5)."""
    from fos_b200 import problems
    lit = _literals()["testPSD"]
    v = problems.svec(np.array(lit["ys"]))
    H = fos.Handle(0)
    got = problems.smat(H.prox_cone("SDP", v))
    np.testing.assert_allclose(got, np.array(lit["projection"]), rtol=1e-10, atol=1e-14)


# ---------------------------------------------------------------------------------------------
# frozen oracle trajectories
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", FIXTURES)
def test_c_oracle_regenerates_golden(oracle, name):
    """The committed trajectories are what oracle/fos_oracle.c produces today (free-running from the
    initial value): a change of the restatement cannot slip through unnoticed."""
    z, P, alg = _load(name)
    O = oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=_direct(z))
    O.set_algorithm(*ALG_SETUPS[alg][0])
    O.set_iterate(O.initial_value())
    n_iter, checki, eps = int(z["n_iter"]), int(z["checki"]), float(z["eps"])
    k = 0
    for i in range(1, n_iter + 1):
        np.testing.assert_allclose(O.get_state("x"), z["before_x"][i - 1], rtol=1e-12, atol=1e-300)
        assert O.s1_calls == z["before_s1_calls"][i - 1]
        out = O.run(i, 1, checki=checki, eps=eps)
        assert O.cgiter == z["cgiter"][i - 1]
        np.testing.assert_allclose(O.get_state("x"), z["after_x"][i - 1], rtol=1e-12, atol=1e-300)
        if i % checki == 0:
            h = out["history"]
            got = [h["i"][0], h["p"][0], h["d"][0], h["g"][0], h["ctx"][0], h["bty"][0], h["kappa"][0], h["tau"][0],
                   h["cgiter"][0], h["status"][0]]
            np.testing.assert_allclose(got, z["records"][k], rtol=1e-10, atol=1e-300, equal_nan=True)
            k += 1
    assert k == len(z["records"])


@pytest.mark.parametrize("name", FIXTURES)
def test_numpy_restatement_lockstep_on_golden(name):
    """The independent NumPy restatement, started from the stored state of every iteration, lands on the
    stored result: the fixtures are well-conditioned enough for the 1e-10 lock-step bar (what the GPU test
    below relies on) and do not encode an accident of the C code."""
    from oracle import np_oracle as npo
    z, P, alg = _load(name)
    M = npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=_direct(z))
    M.set_algorithm(*ALG_SETUPS[alg][0])
    M.checki, M.eps = 10 ** 6, float(z["eps"])
    worst = 0.0
    for i in range(1, int(z["n_iter"]) + 1):
        M.x = z["before_x"][i - 1].copy()
        if not _direct(z):
            if z["before_s1_calls"][i - 1] > 1:
                M.S1.xinit = z["before_xinit"][i - 1].copy()
            M.S1.i = int(z["before_s1_calls"][i - 1])
        M.alpha12, M.t = float(z["before_alpha12"][i - 1]), float(z["before_fista_t"][i - 1])
        for attr, key in (("y", "before_fista_y"), ("p", "before_dykstra_p"), ("q", "before_dykstra_q")):
            setattr(M, attr, z[key][i - 1].copy() if key in z.files else np.zeros(M.N))
        M.i = i
        M.step()
        if not _direct(z):
            assert M.S1.cgiter == z["cgiter"][i - 1], f"iteration {i}: CG count"
        worst = max(worst, rel_err(M.x, z["after_x"][i - 1]))
    tol = _tol(z, alg)
    assert worst < tol, f"{name}: NumPy restatement deviates by {worst:.2e}"


def _sync_state_from_golden(H, z, k, alg):
    H.set_state("x", z["before_x"][k])
    if z["before_s1_calls"][k] > 1:
        H.set_state("xinit", z["before_xinit"][k])
    H.set_info("s1_calls", int(z["before_s1_calls"][k]))
    if alg.startswith("GAPA"):
        H.set_info("alpha12", float(z["before_alpha12"][k]))
    if alg == "FISTA":
        H.set_state("fista_y", z["before_fista_y"][k] if "before_fista_y" in z.files else np.zeros_like(z["before_x"][k]))
        H.set_info("fista_t", float(z["before_fista_t"][k]))
    if alg == "Dykstra":
        for key in ("dykstra_p", "dykstra_q"):
            H.set_state(key, z["before_" + key][k] if "before_" + key in z.files else np.zeros_like(z["before_x"][k]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_lockstep_on_golden(fos, name):
    """CUDA path through the C ABI against the committed fixtures: from the stored state before iteration
    i, ONE iteration gives the stored iterate / relaxed S1 output to 1e-10, the stored CG iteration count,
    S1 call counter, GAPA angle and p/d/g/ctx/bty/kappa/tau record."""
    z, P, alg = _load(name)
    H = load_conic(fos, P)
    if _direct(z):
        H.set_direct(True)
    H.set_algorithm(ALG_SETUPS[alg][1](fos))
    n_iter, checki, eps = int(z["n_iter"]), int(z["checki"]), float(z["eps"])
    tol = _tol(z, alg)
    H.ck(H.L.fos_begin_solve(H.h))
    k = 0
    worst = 0.0
    for i in range(1, n_iter + 1):
        _sync_state_from_golden(H, z, i - 1, alg)
        done, st, rec, _ = H.run(i, 1, checki, eps)
        assert done == 1
        assert H.info("cgiter") == z["cgiter"][i - 1], f"iteration {i}: CG count"
        assert H.info("s1_calls") == z["after_s1_calls"][i - 1]
        e = rel_err(H.get_iterate(), z["after_x"][i - 1])
        worst = max(worst, e)
        assert e < tol, f"iteration {i}: iterate differs by {e:.3e}"
        if alg not in ("FISTA", "Dykstra") and not _direct(z):
            assert rel_err(H.get_state("tmp1"), z["after_tmp1"][i - 1]) < tol
        if alg.startswith("GAPA"):
            assert abs(H.info("alpha12") - z["after_alpha12"][i - 1]) < 1e-9
        if i % checki == 0:
            want = z["records"][k]
            assert len(rec) == 1 and rec[0, 0] == want[0] == i
            rt, at = (1e-8, 1e-11) if _direct(z) else (1e-9, 1e-12)   # as the live-oracle tests of tests/test_gpu_solvers.py
            np.testing.assert_allclose(rec[0, 1:8], want[1:8], rtol=rt, atol=at, equal_nan=True)
            assert rec[0, 8] == want[8] and rec[0, 9] == want[9]
            k += 1
        else:
            assert len(rec) == 0
    print(f"{name}: worst one-step relative deviation {worst:.2e}")


# ---------------------------------------------------------------------------------------------
# Feasibility form (Feasibility.jl, FeasibilityStatus.jl): frozen oracle trajectories
# ---------------------------------------------------------------------------------------------
FEAS_FIXTURES = sorted(p.name for p in GOLDEN.glob("feasibility_*.npz"))


def _load_feas(name):
    z = np.load(GOLDEN / name, allow_pickle=False)
    shape = tuple(int(v) for v in z["A_shape"])
    A = sp.csc_matrix((z["A_data"], z["A_indices"], z["A_indptr"]), shape=shape).toarray()
    cones = [(n, k) for n, k in json.loads(str(z["cones"]))]
    return z, A, z["b"], z["q"], int(z["beta"]), cones, str(z["alg"])


def test_feasibility_fixtures_are_present():
    assert len(FEAS_FIXTURES) >= 4


@pytest.mark.parametrize("name", FEAS_FIXTURES)
def test_c_oracle_regenerates_feasibility_golden(oracle, name):
    z, A, b, q, beta, cones, alg = _load_feas(name)
    O = oracle.OracleFeasibility(A, b, q, beta, cones)
    O.set_algorithm(*ALG_SETUPS[alg][0])
    O.set_iterate(O.initial_value())
    checki, k = int(z["checki"]), 0
    for i in range(1, int(z["n_iter"]) + 1):
        np.testing.assert_allclose(O.get_state("x"), z["before_x"][i - 1], rtol=1e-12, atol=1e-300)
        out = O.run(i, 1, checki=checki, eps=float(z["eps"]))
        assert O.cgiter == z["cgiter"][i - 1]
        np.testing.assert_allclose(O.get_state("x"), z["after_x"][i - 1], rtol=1e-12, atol=1e-300)
        if i % checki == 0:
            h = out["history"]
            np.testing.assert_allclose([h["i"][0], h["p"][0], h["status"][0]], z["records"][k], rtol=1e-10, equal_nan=True)
            k += 1
    assert k == len(z["records"])


@pytest.mark.parametrize("name", FEAS_FIXTURES)
def test_numpy_restatement_lockstep_on_feasibility_golden(name):
    from oracle import np_oracle as npo
    z, A, b, q, beta, cones, alg = _load_feas(name)
    M = npo.NPModel.feasibility(A, b, q, beta, cones)
    M.set_algorithm(*ALG_SETUPS[alg][0])
    M.checki, M.eps = 10 ** 6, float(z["eps"])
    worst = 0.0
    for i in range(1, int(z["n_iter"]) + 1):
        M.x = z["before_x"][i - 1].copy()
        if z["before_s1_calls"][i - 1] > 1:
            M.S1.xinit = z["before_xinit"][i - 1].copy()
        M.S1.i = int(z["before_s1_calls"][i - 1])
        M.alpha12, M.t = float(z["before_alpha12"][i - 1]), float(z["before_fista_t"][i - 1])
        for attr, key in (("y", "before_fista_y"), ("p", "before_dykstra_p"), ("q", "before_dykstra_q")):
            setattr(M, attr, z[key][i - 1].copy() if key in z.files else np.zeros(M.N))
        M.i = i
        M.step()
        assert M.S1.cgiter == z["cgiter"][i - 1], f"iteration {i}: CG count"
        worst = max(worst, rel_err(M.x, z["after_x"][i - 1]))
    assert worst < STEP_TOL, f"{name}: NumPy restatement deviates by {worst:.2e}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", FEAS_FIXTURES)
def test_gpu_lockstep_on_feasibility_golden(fos, name):
    """fos_load_affine_csc + one iteration from the stored state: iterate to 1e-10, CG count, err record
    (FeasibilityStatus.jl:32-72)."""
    from helpers import load_affine
    z, A, b, q, beta, cones, alg = _load_feas(name)
    H = load_affine(fos, A, b, q, beta, cones)
    H.set_algorithm(ALG_SETUPS[alg][1](fos))
    H.ck(H.L.fos_begin_solve(H.h))
    checki, k = int(z["checki"]), 0
    for i in range(1, int(z["n_iter"]) + 1):
        _sync_state_from_golden(H, z, i - 1, alg)
        done, st, rec, _ = H.run(i, 1, checki, float(z["eps"]))
        assert H.info("cgiter") == z["cgiter"][i - 1], f"iteration {i}: CG count"
        e = rel_err(H.get_iterate(), z["after_x"][i - 1])
        assert e < STEP_TOL, f"iteration {i}: iterate differs by {e:.3e}"
        if i % checki == 0:
            np.testing.assert_allclose(rec[0, 1], z["records"][k][1], rtol=1e-9, equal_nan=True)  # err
            k += 1
