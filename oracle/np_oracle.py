"""Second, independent CPU restatement of the hot path in NumPy/SciPy.

TEST INFRASTRUCTURE ONLY (same rules as ``fos_oracle.c``).  Its purpose is to cross-check
the C oracle: two restatements written separately from the same reference lines must agree
to rounding, which is the strongest pin available while the Julia original cannot run here
("parity unpinned" for seeded goldens -- see ``fos_oracle.c``).  It uses LAPACK
(``numpy.linalg.eigh``) for the PSD cone like the reference's ``dspev`` path, whereas the
C oracle uses its own Jacobi sweep.

Citations are to /root/reference/src.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

EPS = np.finfo(np.float64).eps


# ---------------------------------------------------------------------------------------
# cones (cones.jl; arithmetic from ProximalOperators.jl, un-vendored)
# ---------------------------------------------------------------------------------------
def sdp_dim(length: int) -> int:
    d = int(round(math.sqrt(0.25 + 2 * length) - 0.5))
    assert d * (d + 1) // 2 == length
    return d


def svec_to_mat(x):
    """packed lower triangle, column-major -> full symmetric (no scaling applied)."""
    d = sdp_dim(x.size)
    S = np.zeros((d, d))
    k = 0
    for j in range(d):
        S[j:, j] = x[k:k + d - j]
        S[j, j:] = x[k:k + d - j]
        k += d - j
    return S


def mat_to_svec(S):
    d = S.shape[0]
    return np.concatenate([S[j:, j] for j in range(d)])


def prox_sdp(x):
    """IndPSD(scaling=true): diag*=sqrt2, eigh, clamp, repack, diag/=sqrt2."""
    d = sdp_dim(x.size)
    S = svec_to_mat(x)
    S[np.diag_indices(d)] *= math.sqrt(2.0)
    w, V = np.linalg.eigh(S)
    P = (V * np.maximum(w, 0.0)) @ V.T
    P = 0.5 * (P + P.T)
    P[np.diag_indices(d)] /= math.sqrt(2.0)
    return mat_to_svec(P)


def prox_soc(x):
    nx = np.linalg.norm(x[1:])
    t = x[0]
    if t <= -nx:
        return np.zeros_like(x)
    if t >= nx:
        return x.copy()
    r = 0.5 * (1.0 + t / nx)
    y = r * x
    y[0] = r * nx
    return y


def prox_cone(name, x):
    if name == "Free":
        return x.copy()
    if name == "Zero":
        return np.zeros_like(x)
    if name == "NonNeg":
        return np.maximum(x, 0.0)
    if name == "NonPos":
        return np.minimum(x, 0.0)
    if name == "SOC":
        return prox_soc(x)
    if name == "SDP":
        return prox_sdp(x)
    raise NotImplementedError(name)


def prox_cone_dual(name, x):
    """proxDual! (cones.jl:80-85, shortcuts :97-102)."""
    if name == "Zero":
        return prox_cone("Free", x)
    if name == "Free":
        return prox_cone("Zero", x)
    if name in ("NonNeg", "NonPos"):
        return prox_cone(name, x)
    return x + prox_cone(name, -x)


def coneprod_prox(cones, x, dual=False):
    y = np.empty_like(x)
    off = 0
    for name, ln in cones:
        seg = x[off:off + ln]
        y[off:off + ln] = prox_cone_dual(name, seg) if dual else prox_cone(name, seg)
        off += ln
    assert off == x.size
    return y


# ---------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------
class HSDEQ:
    """HSDEMatrixQ (problemforms/HSDE/HSDEAffine.jl:2-65)."""

    def __init__(self, A, b, c):
        self.A = sp.csc_matrix(A)
        self.At = self.A.T.tocsc()
        self.b, self.c = b, c
        self.m, self.n = self.A.shape
        self.am = self.an = self.m + self.n + 1

    def mul(self, B):
        n, m = self.n, self.m
        b1, b2, b3 = B[:n], B[n:n + m], B[n + m]
        y1 = self.At @ b2 + b3 * self.c
        y2 = -(self.A @ b1 - b3 * self.b)
        return np.concatenate([y1, y2, [-(self.c @ b1) - (self.b @ b2)]])

    def mul_t(self, B):
        return -self.mul(B)

    def dense(self):
        A = self.A.toarray()
        n, m = self.n, self.m
        Q = np.zeros((n + m + 1, n + m + 1))
        Q[:n, n:n + m] = A.T
        Q[:n, -1] = self.c
        Q[n:n + m, :n] = -A
        Q[n:n + m, -1] = self.b
        Q[-1, :n] = -self.c
        Q[-1, n:n + m] = -self.b
        return Q


class PlainOp:
    def __init__(self, A):
        self.A = sp.csc_matrix(A)
        self.At = self.A.T.tocsc()
        self.am, self.an = self.A.shape

    def mul(self, x):
        return self.A @ x

    def mul_t(self, x):
        return self.At @ x


def kkt_mul(op, x):
    """KKTMatrix mul! (utilities/affinepluslinear.jl:37-49)."""
    an = op.an
    x1, x2 = x[:an], x[an:]
    return np.concatenate([op.mul_t(x2) + x1, op.mul(x1) - x2])


def conjgrad(x, mul, b, tol, max_iters):
    """conjugategradient! (utilities/conjugategradients.jl:31-55); x updated in place."""
    r = b - mul(x)
    p = r.copy()
    rn = r @ r
    it = 1
    while True:
        Ap = mul(p)
        alpha = rn / (Ap @ p)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= tol or it >= max_iters:
            break
        rnold = rn
        rn = r @ r
        beta = rn / rnold
        p *= beta
        p += r
        it += 1
    return it


class AffinePlusLinear:
    """utilities/affinepluslinear.jl:58-126."""

    def __init__(self, op, b, q, beta, decreasing_accuracy=False):
        self.op, self.beta = op, beta
        self.b = np.zeros(op.am) if b is None else np.asarray(b, float)
        self.q = np.zeros(op.an) if q is None else np.asarray(q, float)
        self.decreasing = decreasing_accuracy
        self.i = 1
        self.cgiter = 0
        self.xinit = None

    def prox(self, x):
        an = self.op.an
        x1, x2 = x[:an], x[an:]
        rhs = np.concatenate([self.beta * self.op.mul_t(x2) + x1 - self.q, self.b])
        if self.xinit is None:
            self.xinit = x.copy()
        y = self.xinit.copy()
        tol = max(0.2 ** math.sqrt(self.i), an * EPS) if self.decreasing else an * EPS
        self.i += 1
        self.cgiter = conjgrad(y, lambda v: kkt_mul(self.op, v), rhs, tol, 1000)
        self.xinit = y.copy()
        y[an:] *= self.beta
        return y


class IndAffineDirect:
    """direct = true (HSDE.jl:10-15): S1 = IndAffine([Q -I], 0), exact projection z - B'(B B')^-1 B z through a
    dense LAPACK solve (independent of the C oracle's hand-written Cholesky)."""

    def __init__(self, op):
        self.op = op
        l = op.an
        Qd = np.column_stack([op.mul(e) for e in np.eye(l)])
        self.B = np.hstack([Qd, -np.eye(l)])
        self.G = self.B @ self.B.T
        self.i, self.cgiter, self.xinit = 1, 0, None

    def prox(self, x):
        return x - self.B.T @ np.linalg.solve(self.G, self.B @ x)


# ---------------------------------------------------------------------------------------
# models + algorithms
# ---------------------------------------------------------------------------------------
class NPModel:
    def __init__(self, form, S1, N, cones1, cones2=None, A=None, b=None, c=None):
        self.form, self.S1, self.N = form, S1, N
        self.K1, self.K2 = cones1, cones2
        self.A, self.b, self.c = A, b, c
        self.alg = ("GAP", 0.8, 1.8, 1.8, 0.0, 100)
        self.alpha12, self.t = 2.0, 1.0
        self.x = np.zeros(N)
        self.y = np.zeros(N)
        self.xold = np.zeros(N)
        self.p = np.zeros(N)
        self.q = np.zeros(N)
        self.status, self.checked = "Continue", False
        self.prev = np.full(N, np.nan)
        self.hist = []
        self.i = 0
        self.lsinterval = 0  # LineSearchWrapper (wrappers/linesearch.jl); 0 = no wrapper
        self.alphabest = None

    @classmethod
    def conic(cls, c, A, b, constr_cones, var_cones, direct=False):
        A = sp.csc_matrix(A)
        m, n = A.shape
        Q = HSDEQ(A, np.asarray(b, float), np.asarray(c, float))
        if direct:
            S1 = IndAffineDirect(Q)  # HSDE.jl:10-15
        else:
            S1 = AffinePlusLinear(Q, None, None, 1, decreasing_accuracy=True)  # HSDE.jl:22
        M = cls(0, S1, 2 * (m + n + 1), constr_cones, var_cones, A, np.asarray(b, float), np.asarray(c, float))
        M.m, M.n = m, n
        M.x[m + n] = 1.0
        M.x[2 * (m + n) + 1] = 1.0
        return M

    @classmethod
    def feasibility(cls, A, b, q, beta, cones, decreasing_accuracy=False):
        op = PlainOp(A)
        S1 = AffinePlusLinear(op, b, q, beta, decreasing_accuracy)
        return cls(1, S1, op.am + op.an, cones)

    def set_algorithm(self, name, alpha=0.8, alpha1=1.8, alpha2=1.8, beta=0.0, iproj=100):
        self.alg = (name, alpha, alpha1, alpha2, beta, iproj)
        self.alpha12, self.t = 2.0, 1.0
        self.p[:] = 0
        self.q[:] = 0

    def P1(self, x):
        return self.S1.prox(x)

    def P2(self, x):
        if self.form == 1:
            return coneprod_prox(self.K1, x)
        m, n = self.m, self.n
        nu = n + m + 1
        y = np.empty_like(x)
        y[:n] = coneprod_prox(self.K2, x[:n])
        y[n:n + m] = coneprod_prox(self.K1, x[n:n + m], dual=True)
        y[nu - 1] = max(x[nu - 1], 0.0)
        y[nu:nu + n] = coneprod_prox(self.K2, x[nu:nu + n], dual=True)
        y[nu + n:nu + n + m] = coneprod_prox(self.K1, x[nu + n:nu + n + m])
        y[2 * nu - 1] = max(x[2 * nu - 1], 0.0)
        return y

    def check(self, z, override=False):
        if self.form == 1:  # FeasibilityStatus.jl:32-72
            if self.i % self.checki == 0 or override:
                err = np.linalg.norm(self.prev - z)
                st = "Optimal" if err <= self.eps else "Continue"
                self.hist.append(dict(i=self.i, err=err, cgiter=self.S1.cgiter, status=st))
                self.status, self.checked = st, True
            else:
                self.checked = False
            self.prev = z.copy()
            return
        if not (self.i % self.checki == 0 or override):  # HSDEStatus.jl:27-71
            self.checked = False
            return
        m, n, A, b, c, eps = self.m, self.n, self.A, self.b, self.c, self.eps
        nu = n + m + 1
        x, y, r, s = z[:n], z[n:n + m], z[nu:nu + n], z[nu + n:nu + n + m]
        tau, kap = z[nu - 1], z[2 * nu - 1]
        nb, nc = np.linalg.norm(b), np.linalg.norm(c)
        with np.errstate(all="ignore"):
            Ax, Aty = A @ x, A.T @ y
            p = np.linalg.norm(Ax / tau + s / tau - b) / abs(1 + nb)
            d = np.linalg.norm(Aty / tau + c - r / tau) / abs(1 + nc)
            ctx, bty = c @ x, b @ y
            g = abs(ctx / tau + bty / tau) / (1 + abs(ctx / tau) + abs(bty / tau))
            st = "Continue"
            if p <= eps * (1 + nb) and d <= eps * (1 + nc) and g <= eps * (1 + abs(ctx / tau) + abs(bty / tau)):
                st = "Optimal"
            elif np.linalg.norm(Ax + s) <= eps * (np.float64(-ctx) / np.float64(nc)):
                st = "Unbounded"
            elif np.linalg.norm(Aty) <= eps * (np.float64(-bty) / np.float64(nb)):
                st = "Infeasible"
        self.hist.append(dict(i=self.i, p=p, d=d, g=g, ctx=ctx, bty=bty, kappa=kap, tau=tau,
                              cgiter=self.S1.cgiter, status=st))
        self.status, self.checked = st, True

    def _ls_step(self, a1, a2):
        """step(::LineSearchWrapper, ...) on a line-search iteration (wrappers/linesearch.jl:42-72)."""
        def S1(v):
            return a1 * self.P1(v) + (1 - a1) * v

        x0 = self.x.copy()
        t2 = S1(x0)
        y = self.P2(t2)
        self.check(y)
        xn = a2 * y + (1 - a2) * t2
        res = xn - x0
        best, abest, al = math.inf, 1.0, 0.1
        for _ in range(31):
            al = al * 1.8
            xt = x0 + al * res
            t2 = S1(xt)
            t3 = self.P2(t2)
            t3 = a2 * t3 + (1 - a2) * t2
            d = xt - t3
            testres = math.sqrt(float(d @ d))
            if testres < best:
                best, abest = testres, al
        self.alphabest = abest
        self.x = x0 + abest * res

    def step(self):
        name, a, a1, a2, bt, iproj = self.alg
        x = self.x
        if name in ("GAP", "GAPA") and self.lsinterval > 0 and self.i % self.lsinterval == 0:
            if name == "GAPA":
                a1 = a2 = self.alpha12
            self._ls_step(a1, a2)
            return
        if name in ("GAP", "GAPA"):
            if name == "GAPA":
                a1 = a2 = self.alpha12
            t1 = self.P1(x)
            t1 = a1 * t1 + (1 - a1) * x
            self.tmp1 = t1  # gap.jl:48 / gapa.jl:74: the relaxed S1 output (data.tmp1)
            t2 = self.P2(t1)
            self.check(t2)
            t2 = a2 * t2 + (1 - a2) * t1
            if name == "GAPA":  # gapa.jl:96-101
                d1, d2 = t2 - t1, t1 - x
                with np.errstate(all="ignore"):
                    scl = abs(d1 @ d2) / math.sqrt((d1 @ d1) * (d2 @ d2)) if (d1 @ d1) * (d2 @ d2) != 0 else float("nan")
                scl = 0.0 if math.isnan(scl) else min(max(scl, 0.0), 1.0)
                self.alpha12 = (1 - bt) * (2 / (1 + math.sqrt(1 - scl * scl))) + bt * 2.0
            self.x = a * t2 + (1 - a) * x
        elif name == "FISTA":  # fista.jl:28-48
            if self.i == 1:
                self.y = x.copy()
            y = self.y
            t1 = self.P1(y)
            t1 = a * t1 + (1 - a) * y
            self.xold = x.copy()
            self.x = self.P2(t1)
            self.check(self.x)
            told = self.t
            self.t = (1 + math.sqrt(1 + 4 * told * told)) / 2
            self.y = self.x + (told - 1) / self.t * (self.x - self.xold)
        elif name == "Dykstra":  # dykstra.jl:26-37
            y = self.P1(x + self.p)
            self.p = x + self.p - y
            self.x = self.P2(y + self.q)
            self.check(self.x)
            self.q = y + self.q - self.x
        elif name == "GAPP":  # gapproj.jl:29-74
            t1 = self.P1(x)
            if self.i % iproj == 0:
                t2 = self.P2(t1)
                res = self.P1(t2) - t1
                best, abest = math.inf, -1.0
                for k in range(21):
                    at = 2.0 ** k
                    t3 = t1 + at * res
                    nt = np.linalg.norm(self.P2(t3) - t3)
                    if nt < best:
                        abest, best = at, nt
                t1 = t1 + abest * res
                self.tmp1 = t1
                t2 = self.P2(t1)
                self.check(t2)
                self.x = a2 * t2 + (1 - a2) * t1
            else:
                t1 = a1 * t1 + (1 - a1) * x
                self.tmp1 = t1
                t2 = self.P2(t1)
                self.check(t2)
                t2 = a2 * t2 + (1 - a2) * t1
                self.x = a * t2 + (1 - a) * x
        else:
            raise ValueError(name)

    def solve(self, max_iters=10000, checki=100, eps=1e-5, trace=False):
        """solve!(model) + iterate (solverwrapper.jl:2-41)."""
        self.checki, self.eps = checki, eps
        self.status, self.checked, self.hist = "Continue", False, []
        self.prev = np.full(self.N, np.nan)
        tr = []
        done = 0
        for i in range(1, max_iters + 1):
            self.i = i
            self.step()
            done += 1
            if trace:
                tr.append(self.x.copy())
            if self.status != "Continue":
                break
        guess = self.P2(self.P1(self.x))
        if not self.checked:
            self.check(guess, override=True)
        st = "Indeterminate" if self.status == "Continue" else self.status
        return {"iterations": done, "status": st, "guess": guess, "history": self.hist,
                "trace": np.array(tr) if trace else None}
