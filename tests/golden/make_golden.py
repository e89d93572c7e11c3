"""Generates the committed fixtures of tests/golden/.

Two kinds of fixture, and the difference matters:

* reference_literals.json -- material copied BY VALUE from the reference's own tests (the only
  literal known-answer data they hold): the 2x2 matrix of test/testPSD.jl:3-4 with its projection
  (worked out by hand from the eigen-decomposition, see the "derivation" entry), the printed header /
  row strings of test/testprint.jl:15-19, the constructor defaults of solvers/*.jl, and the optimum
  recorded in test/testDRandGAPA.jl:12,15 (kept for the day a Julia RNG is available; it cannot be
  reproduced here because it depends on Julia's `randn` stream).
* lockstep_[direct_]<problem>_<algorithm>.npz -- trajectories of the CPU oracle (oracle/fos_oracle.c) on
  seeded, well-conditioned instances: the complete solver state before every iteration and the
  oracle's result after it.  THESE ARE ORACLE OUTPUTS, NOT REFERENCE OUTPUTS: the reference is Julia
  and cannot run in this environment.  They freeze the oracle (a change of the restatement shows up as
  a diff of committed data) and give the GPU parity tests inputs/outputs that do not depend on the
  oracle being rebuilt on the GPU box.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_ITER, CHECKI, EPS = 20, 5, 1e-12

# (problem kind, algorithm of helpers.ALG_SETUPS)
CASES = [("nnls", "DR"), ("nnls", "GAP"), ("nnls", "AP"), ("nnls", "GAPA_b"), ("nnls", "FISTA"), ("nnls", "Dykstra"),
         ("nnls", "GAPP"), ("lasso", "DR"), ("sdp", "GAP"), ("socls", "GAPA")]
WELL = {"nnls": 0.02, "lasso": 0.1, "socls": 0.02, "sdp": 1.0}  # as tests/test_gpu_solvers.py
# direct = true (HSDE.jl:10-15) on the UNSCALED instances: no truncated CG, nothing amplifies rounding
DIRECT_CASES = [("nnls", "DR"), ("nnls", "GAPP"), ("sdp", "GAP")]


def build_problem(kind, direct=False):
    from fos_b200 import problems
    if kind == "nnls":
        return problems.nnls_conic(40, 50, seed=1, scale=1.0 if direct else WELL[kind])
    if kind == "lasso":
        return problems.lasso_like(60, 130, seed=2, scale=WELL[kind])
    if kind == "socls":
        return problems.soc_constrained_ls(300, 20, seed=3, scale=WELL[kind])
    if kind == "sdp":
        return problems.sdp_nearest_correlation(6, seed=4)
    raise KeyError(kind)


def snapshot(O):
    return {"x": O.get_state("x"), "xinit": O.get_state("xinit"), "fista_y": O.get_state("fista_y"),
            "dykstra_p": O.get_state("dykstra_p"), "dykstra_q": O.get_state("dykstra_q"),
            "s1_calls": O.s1_calls, "alpha12": O.alpha12, "fista_t": O.fista_t}


def make_lockstep(kind, alg, direct=False):
    from helpers import ALG_SETUPS
    from oracle import fos_oracle
    P = build_problem(kind, direct)
    O = fos_oracle.OracleConic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=direct)
    O.set_algorithm(*ALG_SETUPS[alg][0])
    O.set_iterate(O.initial_value())
    before, after_x, after_tmp1, cgiter, s1_after, a12_after, recs = [], [], [], [], [], [], []
    for i in range(1, N_ITER + 1):
        before.append(snapshot(O))
        out = O.run(i, 1, checki=CHECKI, eps=EPS)
        h = out["history"]
        after_x.append(O.get_state("x"))
        after_tmp1.append(O.get_state("tmp1"))
        cgiter.append(O.cgiter)
        s1_after.append(O.s1_calls)
        a12_after.append(O.alpha12)
        if i % CHECKI == 0:
            assert len(h["i"]) == 1
            recs.append([h["i"][0], h["p"][0], h["d"][0], h["g"][0], h["ctx"][0], h["bty"][0], h["kappa"][0],
                         h["tau"][0], h["cgiter"][0], h["status"][0]])
        else:
            assert len(h["i"]) == 0
    A = sp.csc_matrix(P.A)
    A.sort_indices()
    data = {
        "c": np.asarray(P.c, float), "b": np.asarray(P.b, float),
        "A_data": A.data.astype(float), "A_indices": A.indices.astype(np.int64), "A_indptr": A.indptr.astype(np.int64),
        "A_shape": np.array(A.shape, dtype=np.int64),
        "cones": np.array(json.dumps({"constr": [[n, int(k)] for n, k in P.constr_cones],
                                      "var": [[n, int(k)] for n, k in P.var_cones]})),
        "alg": np.array(alg), "direct": np.int64(1 if direct else 0), "n_iter": np.int64(N_ITER), "checki": np.int64(CHECKI), "eps": np.float64(EPS),
        "after_x": np.array(after_x), "after_tmp1": np.array(after_tmp1), "cgiter": np.array(cgiter, dtype=np.int64),
        "after_s1_calls": np.array(s1_after, dtype=np.int64), "after_alpha12": np.array(a12_after, float),
        "records": np.array(recs, float),
    }
    for key in ("x", "xinit", "fista_y", "dykstra_p", "dykstra_q"):
        arr = np.array([s[key] for s in before])
        if key in ("x", "xinit") or np.any(arr != 0.0):
            data["before_" + key] = arr
    data["before_s1_calls"] = np.array([s["s1_calls"] for s in before], dtype=np.int64)
    data["before_alpha12"] = np.array([s["alpha12"] for s in before], float)
    data["before_fista_t"] = np.array([s["fista_t"] for s in before], float)
    path = HERE / (f"lockstep_direct_{kind}_{alg}.npz" if direct else f"lockstep_{kind}_{alg}.npz")
    np.savez_compressed(path, **data)
    return path


FEAS_CASES = ["DR", "GAPA", "FISTA", "Dykstra"]


def make_feasibility(alg):
    """Feasibility form (Feasibility.jl:2-6): find x in {A x = b} and x >= 0, posed as AffinePlusLinear(A, b, 0, 1)
    on [x; z] with z in Zero (test/testfeasibility.jl:5-19 shape); instance of tests/test_gpu_solvers.py::
    test_feasibility_lockstep."""
    from helpers import ALG_SETUPS
    from fos_b200 import problems
    from oracle import fos_oracle
    A, b, cones = problems.feasibility_problem(30, 70, seed=5)
    A = A * 0.05
    b = b * 0.05
    q = np.zeros(70)
    O = fos_oracle.OracleFeasibility(A, b, q, 1, cones)
    O.set_algorithm(*ALG_SETUPS[alg][0])
    O.set_iterate(O.initial_value())
    n_iter, checki = 21, 7
    before, after_x, cgiter, errs = [], [], [], []
    for i in range(1, n_iter + 1):
        before.append(snapshot(O))
        out = O.run(i, 1, checki=checki, eps=EPS)
        after_x.append(O.get_state("x"))
        cgiter.append(O.cgiter)
        if i % checki == 0:
            errs.append([out["history"]["i"][0], out["history"]["p"][0], out["history"]["status"][0]])
    Ac = sp.csc_matrix(A)
    Ac.sort_indices()
    data = {"form": np.array("feasibility"), "alg": np.array(alg), "b": np.asarray(b, float), "q": q, "beta": np.int64(1),
            "A_data": Ac.data.astype(float), "A_indices": Ac.indices.astype(np.int64), "A_indptr": Ac.indptr.astype(np.int64),
            "A_shape": np.array(Ac.shape, dtype=np.int64),
            "cones": np.array(json.dumps([[n, int(k)] for n, k in cones])),
            "n_iter": np.int64(n_iter), "checki": np.int64(checki), "eps": np.float64(EPS),
            "after_x": np.array(after_x), "cgiter": np.array(cgiter, dtype=np.int64), "records": np.array(errs, float)}
    for key in ("x", "xinit", "fista_y", "dykstra_p", "dykstra_q"):
        arr = np.array([s_[key] for s_ in before])
        if key in ("x", "xinit") or np.any(arr != 0.0):
            data["before_" + key] = arr
    data["before_s1_calls"] = np.array([s_["s1_calls"] for s_ in before], dtype=np.int64)
    data["before_alpha12"] = np.array([s_["alpha12"] for s_ in before], float)
    data["before_fista_t"] = np.array([s_["fista_t"] for s_ in before], float)
    path = HERE / f"feasibility_{alg}.npz"
    np.savez_compressed(path, **data)
    return path


def make_literals():
    # test/testPSD.jl:3-4: ys = [-0.0064709 -0.22443; -0.22443 -1.02411] has the eigenvalues -1.0714074874680
    # and +0.0408265874680; its projection onto the PSD cone (what IndPSD returns and what SCS returns for the
    # same problem, testPSD.jl:14-19) is lambda_+ v_+ v_+' -- the numbers below, also quoted in BASELINE.md
    lit = {
        "source": "literal values of /root/reference/test and /root/reference/src, copied by value",
        "testPSD": {"cite": "test/testPSD.jl:3-4,14-25", "ys": [[-0.0064709, -0.22443], [-0.22443, -1.02411]],
                    "eigenvalues": [-1.071407487468016, 0.0408265874680158],
                    "projection": [[0.03909044662082823, -0.00823811392936668],
                                   [-0.00823811392936668, 0.00173614084718757]],
                    "derivation": "closed form for a symmetric 2x2 matrix: keep the positive eigenpair",
                    "tolerance_DR_vs_projection": 1e-8},
        "testprint": {"cite": "test/testprint.jl:15-19",
                      "header_with_cg": " Iter | pri res | dua res | rel gap | pri obj | dua obj | kap/tau | cg  | time",
                      "header_direct": " Iter | pri res | dua res | rel gap | pri obj | dua obj | kap/tau | time",
                      "row_prefixes": ["   100|", "   200|"], "found": "Found solution i=200"},
        "testDRandGAPA": {"cite": "test/testDRandGAPA.jl:12,15 (needs Julia's randn stream: not reproducible here)",
                          "optval_julia_ge_1_5": 10.945929126466417, "optval_julia_lt_1_5": 12.38418747141913},
        "defaults": {"cite": "solvers/gap.jl:14-22, gapa.jl:20-21, fista.jl:9, dykstra.jl:9, gapproj.jl:14",
                     "GAP": {"alpha": 0.8, "alpha1": 1.8, "alpha2": 1.8, "direct": False},
                     "DR": {"alpha": 0.5, "alpha1": 2.0, "alpha2": 2.0},
                     "AP": {"alpha": 1.0, "alpha1": 1.0, "alpha2": 1.0},
                     "GAPA": {"alpha": 1.0, "beta": 0.0, "direct": False},
                     "FISTA": {"alpha": 1.0, "direct": False},
                     "Dykstra": {"direct": False},
                     "GAPP": {"alpha": 0.8, "alpha1": 1.8, "alpha2": 1.8, "iproj": 100, "direct": True}},
        "solve_defaults": {"cite": "FOSSolverInterface.jl / solverwrapper.jl kwargs",
                           "max_iters": 10000, "checki": 100, "eps": 1e-5},
    }
    (HERE / "reference_literals.json").write_text(json.dumps(lit, indent=1) + "\n")


if __name__ == "__main__":
    make_literals()
    total = 0
    for direct, cases in ((False, CASES), (True, DIRECT_CASES)):
        for kind, alg in cases:
            p = make_lockstep(kind, alg, direct)
            total += p.stat().st_size
            print(p.name, p.stat().st_size)
    for alg in FEAS_CASES:
        p = make_feasibility(alg)
        total += p.stat().st_size
        print(p.name, p.stat().st_size)
    print("total bytes", total)
