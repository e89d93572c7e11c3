"""A NumPy-backed stand-in for the library handle (`fos.Handle`), used ONLY by tests/test_harness_selfcheck.py to
run the bodies of the GPU parity tests on a machine without a GPU: it checks the TESTS (fixture keys, record
columns, state plumbing, tolerances that any correct implementation meets), not the product.  It wraps
oracle/np_oracle.py, the independent restatement, so the comparisons against the C oracle and against the
committed fixtures stay meaningful.  Nothing under firstordersolvers.jl_b200/ imports this."""
import types

import numpy as np

from oracle import np_oracle as npo

ALG_NAMES = {0: "GAP", 1: "GAPA", 2: "FISTA", 3: "Dykstra", 4: "GAPP"}
STATUS_CODES = {"Continue": 0, "Optimal": 1, "Unbounded": 2, "Infeasible": 3, "Indeterminate": 4}


class _FakeLib:
    def fos_begin_solve(self, h):
        m = h.M
        m.status, m.checked, m.hist = "Continue", False, []
        m.prev = np.full(m.N, np.nan)
        return 0


class StandInHandle:
    def __init__(self, build):
        self._build = build           # direct -> NPModel
        self.M = build(False)
        self.L = _FakeLib()
        self.h = self
        self._alg = None
        self._warned = False

    def ck(self, rc):
        assert rc == 0

    def set_algorithm(self, alg):
        code, a, a1, a2, b, ip = alg._params()
        self._alg = (ALG_NAMES[code], a, a1, a2, b, ip)
        for m in ([self.M] if self.M is not None else []) + getattr(self, "Ms", []):
            m.set_algorithm(*self._alg)

    # batch mode: B independent models
    def load_conic_batch(self, A, b, c, constr_cones, var_cones, device_ptr=None):
        self.Ms = [npo.NPModel.conic(c[j], A[j], b[j], constr_cones, var_cones) for j in range(len(A))]

    def solve_batch(self, max_iters, checki, eps):
        done, st, recs, guess = [], [], [], []
        for m in self.Ms:
            r = m.solve(max_iters=max_iters, checki=checki, eps=eps)
            self.M = m
            done.append(r["iterations"])
            st.append(STATUS_CODES[r["status"]])
            recs.append(self._records())
            guess.append(r["guess"])
        self.M = None
        return np.array(done), np.array(st), recs, np.array(guess)

    def set_direct(self, on=True):
        self.M = self._build(bool(on))
        if self._alg:
            self.M.set_algorithm(*self._alg)

    def set_initial_iterate(self):
        pass                          # NPModel starts from the initial value

    def set_state(self, which, z):
        z = np.array(z, float)
        if which == "x":
            self.M.x = z
        elif which == "xinit":
            self.M.S1.xinit = z
        elif which == "fista_y":
            self.M.y = z
        elif which == "dykstra_p":
            self.M.p = z
        elif which == "dykstra_q":
            self.M.q = z
        else:
            raise KeyError(which)

    def get_state(self, which):
        return {"x": self.M.x, "tmp1": getattr(self.M, "tmp1", None)}[which].copy()

    def get_iterate(self):
        return self.M.x.copy()

    def set_info(self, which, v):
        if which == "s1_calls":
            self.M.S1.i = int(v)
        elif which == "alpha12":
            self.M.alpha12 = float(v)
        elif which == "fista_t":
            self.M.t = float(v)
        else:
            raise KeyError(which)

    def info(self, which):
        if which == "cgiter":
            return self.M.S1.cgiter
        if which == "s1_calls":
            return self.M.S1.i
        if which == "alpha12":
            return self.M.alpha12
        if which == "cg_warned":
            return 1 if self._warned else 0
        raise KeyError(which)

    def _records(self):
        rows = []
        for r in self.M.hist:
            if "err" in r:
                rows.append([r["i"], r["err"], 0, 0, 0, 0, 0, 0, r["cgiter"], STATUS_CODES[r["status"]]])
            else:
                rows.append([r["i"], r["p"], r["d"], r["g"], r["ctx"], r["bty"], r["kappa"], r["tau"], r["cgiter"],
                             STATUS_CODES[r["status"]]])
        return np.array(rows, float).reshape(-1, 10)

    def run(self, i_start, n_iters, checki, eps, trace=False):
        m = self.M
        m.checki, m.eps, m.hist = checki, eps, []
        done = 0
        for i in range(i_start, i_start + n_iters):
            m.i = i
            m.step()
            done += 1
            if m.status != "Continue":
                break
        return done, STATUS_CODES[m.status], self._records(), None

    def solve(self, max_iters, checki, eps):
        r = self.M.solve(max_iters=max_iters, checki=checki, eps=eps)
        return r["iterations"], STATUS_CODES[r["status"]], self._records(), r["guess"]

    def affine_prox(self, z):
        y = self.M.S1.prox(np.array(z, float))
        self._warned = self._warned or self.M.S1.cgiter == 1000
        return y

    def prox_cone(self, name, x, dual=False):
        return npo.prox_cone_dual(name, np.array(x, float)) if dual else npo.prox_cone(name, np.array(x, float))


def load_conic(fos, P, storage="auto", **options):
    return StandInHandle(lambda direct: npo.NPModel.conic(P.c, P.A, P.b, P.constr_cones, P.var_cones, direct=direct))


def load_affine(fos, A, b, q, beta, cones, decreasing=False, storage="auto", **options):
    return StandInHandle(lambda direct: npo.NPModel.feasibility(A, b, q, beta, cones, decreasing_accuracy=decreasing))


def standin_module(real_fos):
    """A module-like object with the real constructors (GAP, DR, ...) and the stand-in handle."""
    ns = types.SimpleNamespace(**{k: getattr(real_fos, k) for k in dir(real_fos) if not k.startswith("__")})
    ns.Handle = lambda device=0: StandInHandle(lambda direct: None)
    return ns
