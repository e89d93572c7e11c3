"""Synthetic conic-form problem generators for the five BASELINE.json configs.

Conic forms are hand-built in the MathProgBase convention (``b - A ξ ∈ K1``, ``ξ ∈ K2``): the
reference's own route (Convex.jl 0.12 lowering + Julia's RNG) does not exist in this
environment, so seeds and lowering are ours; the GPU path and the oracle are always compared
on the identical ``(c, A, b, cones)``.

Everything here is host-side NumPy and small enough for tests; ``bench.py`` builds the full-size
C2 matrix directly on the device with the same recipe.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


@dataclass
class ConicProblem:
    c: np.ndarray
    A: object  # scipy.sparse matrix or dense ndarray, m x n
    b: np.ndarray
    constr_cones: list  # [(name, length), ...] covering 1:m
    var_cones: list     # covering 1:n
    name: str = ""

    @property
    def m(self):
        return self.A.shape[0]

    @property
    def n(self):
        return self.A.shape[1]


def nnls_conic(rows=40, cols=50, seed=1, scale=1.0) -> ConicProblem:
    """C1 / C5: minimise ||D x - d|| s.t. x >= 0 as  min t  s.t. (t, Dx-d) in SOC, x in NonNeg.
    ξ = (t, x);  m = rows + 1 + cols,  n = cols + 1  (40x50 -> m = 91, n = 51).
    `scale` multiplies D (scale << 1 gives a KKT matrix with eigenvalues near +-1, on which the
    reference's CG does not amplify rounding: used by the strict lock-step parity tests)."""
    rng = np.random.default_rng(seed)
    D = rng.standard_normal((rows, cols)) * scale
    d = rng.standard_normal(rows)
    n = cols + 1
    top = sp.hstack([sp.csr_matrix(([-1.0], ([0], [0])), shape=(1, 1)), sp.csr_matrix((1, cols))])
    mid = sp.hstack([sp.csr_matrix((rows, 1)), sp.csr_matrix(-D)])
    bot = sp.hstack([sp.csr_matrix((cols, 1)), -sp.identity(cols, format="csr")])
    A = sp.vstack([top, mid, bot]).tocsc()
    b = np.concatenate([[0.0], -d, np.zeros(cols)])
    c = np.zeros(n)
    c[0] = 1.0
    return ConicProblem(c, A, b, [("SOC", rows + 1), ("NonNeg", cols)], [("Free", n)], f"nnls{rows}x{cols}")


def _sample_in_cone(rng, cones, dual=False):
    """A point in the product cone (or its dual)."""
    parts = []
    for name, ln in cones:
        v = rng.standard_normal(ln)
        if name == "Zero":
            v = v if dual else np.zeros(ln)
        elif name == "Free":
            v = np.zeros(ln) if dual else v
        elif name == "NonNeg":
            v = np.abs(v)
        elif name == "NonPos":
            v = -np.abs(v)
        elif name == "SOC":
            v[0] = np.linalg.norm(v[1:]) + abs(v[0])
        elif name == "SDP":
            d = int(round(np.sqrt(0.25 + 2 * ln) - 0.5))
            G = rng.standard_normal((d, d))
            S = G @ G.T / d
            S[np.triu_indices(d, 1)] *= np.sqrt(2.0)
            S[np.tril_indices(d, -1)] *= np.sqrt(2.0)
            v = np.concatenate([S[j:, j] for j in range(d)])
        else:
            raise NotImplementedError(name)
        parts.append(v)
    return np.concatenate(parts) if parts else np.zeros(0)


def random_feasible_conic(m, n, constr_cones, seed=2, density=None, dense=True, scale=1.0) -> ConicProblem:
    """C2-style: A = randn(m,n)/sqrt(n); primal and dual strictly feasible by construction
    (b = A ξ* + s*, s* in K1;  c = -A' y*, y* in K1*), variables free."""
    rng = np.random.default_rng(seed)
    if density is None:
        A = rng.standard_normal((m, n)) / np.sqrt(n) * scale
        Aop = A
        A_out = A if dense else sp.csc_matrix(A)
    else:
        A_out = sp.random(m, n, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csc")
        A_out = A_out / np.sqrt(max(density * n, 1.0)) * scale
        Aop = A_out
    xi = rng.standard_normal(n)
    s = _sample_in_cone(rng, constr_cones)
    y = _sample_in_cone(rng, constr_cones, dual=True)
    b = Aop @ xi + s
    c = -(Aop.T @ y)
    return ConicProblem(np.asarray(c).ravel(), A_out, np.asarray(b).ravel(), list(constr_cones), [("Free", n)],
                        f"rand{m}x{n}")


def lasso_like(m=200, n=400, seed=2, dense=True, scale=1.0) -> ConicProblem:
    """C2 at test scale: rows split K1 = Zero(m/2) + NonNeg(m - m/2), DR(0.5)."""
    h = m // 2
    return random_feasible_conic(m, n, [("Zero", h), ("NonNeg", m - h)], seed=seed, dense=dense, scale=scale)


def soc_constrained_ls(md=300, nx=60, seed=3, rho=None, scale=1.0) -> ConicProblem:
    """C3: min t s.t. ||D x - d|| <= t, ||x|| <= rho.  ξ = (t, x); K1 = SOC(md+1) + SOC(nx+1)."""
    rng = np.random.default_rng(seed)
    D = rng.standard_normal((md, nx)) / np.sqrt(nx) * scale
    x0 = rng.standard_normal(nx)
    d = D @ x0 + 0.1 * rng.standard_normal(md)
    if rho is None:
        rho = 0.5 * np.linalg.norm(x0)
    n = nx + 1
    r1 = sp.hstack([sp.csr_matrix(([-1.0], ([0], [0])), shape=(1, 1)), sp.csr_matrix((1, nx))])
    r2 = sp.hstack([sp.csr_matrix((md, 1)), sp.csr_matrix(-D)])
    r3 = sp.csr_matrix((1, n))
    r4 = sp.hstack([sp.csr_matrix((nx, 1)), -sp.identity(nx, format="csr")])
    A = sp.vstack([r1, r2, r3, r4]).tocsc()
    b = np.concatenate([[0.0], -d, [rho], np.zeros(nx)])
    c = np.zeros(n)
    c[0] = 1.0
    return ConicProblem(c, A, b, [("SOC", md + 1), ("SOC", nx + 1)], [("Free", n)], f"socls{md}x{nx}")


def svec(S):
    """symmetric matrix -> packed lower triangle (column-major) with sqrt(2) off-diagonals."""
    d = S.shape[0]
    T = S.astype(float).copy()
    T[np.tril_indices(d, -1)] *= np.sqrt(2.0)
    return np.concatenate([T[j:, j] for j in range(d)])


def smat(v):
    d = int(round(np.sqrt(0.25 + 2 * v.size) - 0.5))
    S = np.zeros((d, d))
    k = 0
    for j in range(d):
        col = v[k:k + d - j].copy()
        col[1:] /= np.sqrt(2.0)
        S[j:, j] = col
        S[j, j:] = col
        k += d - j
    return S


def sdp_nearest_correlation(d=8, seed=4) -> ConicProblem:
    """C4: min <C, X> s.t. diag(X) = 1, X PSD.  ξ = svec(X) free; K1 = Zero(d) + SDP(d(d+1)/2):
    rows 1..d pick the diagonal entries, the remaining block is -I (0 - (-ξ) = ξ in SDP)."""
    rng = np.random.default_rng(seed)
    n = d * (d + 1) // 2
    G = rng.standard_normal((d, d))
    Cm = (G + G.T) / 2
    c = svec(Cm)
    diag_idx = [j * d - j * (j - 1) // 2 for j in range(d)]
    sel = sp.csr_matrix((np.ones(d), (np.arange(d), diag_idx)), shape=(d, n))
    A = sp.vstack([sel, -sp.identity(n, format="csr")]).tocsc()
    b = np.concatenate([np.ones(d), np.zeros(n)])
    return ConicProblem(c, A, b, [("Zero", d), ("SDP", n)], [("Free", n)], f"sdp{d}")


def psd_projection_problem(ys) -> ConicProblem:
    """test/testPSD.jl:16-25 shape: minimise ||vec(Y - ys)|| s.t. Y PSD, as
    min t s.t. (t, svec(Y) - svec(ys)) in SOC, svec(Y) in SDP.  ξ = (t, svec Y)."""
    ys = np.asarray(ys, float)
    d = ys.shape[0]
    k = d * (d + 1) // 2
    n = k + 1
    v = svec((ys + ys.T) / 2)
    r1 = sp.hstack([sp.csr_matrix(([-1.0], ([0], [0])), shape=(1, 1)), sp.csr_matrix((1, k))])
    r2 = sp.hstack([sp.csr_matrix((k, 1)), -sp.identity(k, format="csr")])
    r3 = sp.hstack([sp.csr_matrix((k, 1)), -sp.identity(k, format="csr")])
    A = sp.vstack([r1, r2, r3]).tocsc()
    b = np.concatenate([[0.0], -v, np.zeros(k)])
    c = np.zeros(n)
    c[0] = 1.0
    return ConicProblem(c, A, b, [("SOC", k + 1), ("SDP", k)], [("Free", n)], f"psdproj{d}")


def infeasible_lp(m=30, n=12, seed=7) -> ConicProblem:
    """An LP  min c'x s.t. b - A x >= 0  with a Farkas certificate built in: y > 0 with A'y = 0 and b'y = -1, so no
    x is feasible.  Drives the :Infeasible branch of checkstatus (HSDEStatus.jl:62-63)."""
    rng = np.random.default_rng(seed)
    y = np.abs(rng.standard_normal(m)) + 0.1
    A = rng.standard_normal((m, n))
    A = A - np.outer(y, A.T @ y) / (y @ y)          # A'y = 0
    b = rng.standard_normal(m)
    b = b - y * (b @ y) / (y @ y) - y / (y @ y)       # b'y = -1
    c = rng.standard_normal(n)
    return ConicProblem(c, sp.csc_matrix(A), b, [("NonNeg", m)], [("Free", n)], f"infeasible_lp{m}x{n}")


def unbounded_lp(m=30, n=12, seed=8) -> ConicProblem:
    """An LP  min c'x s.t. b - A x >= 0  that is feasible (x = 0, b > 0) and has a recession direction d with
    A d <= 0 and c'd < 0.  Drives the :Unbounded branch of checkstatus (HSDEStatus.jl:60-61)."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal(n)
    A = rng.standard_normal((m, n))
    A = A * (-np.sign(A @ d))[:, None]               # A d = -|A d| <= 0
    b = np.abs(rng.standard_normal(m)) + 0.5
    c = -d + 0.1 * rng.standard_normal(n)
    if c @ d >= 0:
        c = -d
    return ConicProblem(c, sp.csc_matrix(A), b, [("NonNeg", m)], [("Free", n)], f"unbounded_lp{m}x{n}")


def stiff_feasibility_problem(am=60, an=120, seed=11, decades=3):
    """A Feasibility instance on which the CG of AffinePlusLinear cannot reach its ABSOLUTE tolerance n*eps
    (affinepluslinear.jl:108-112): singular values of A from 1 to 10^decades.  With decades = 3 the solve stops at
    the max_iters = 1000 cap (affinepluslinear.jl:115) and raises the @warn of :120 / conjugategradients.jl:53."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((am, am)))
    V, _ = np.linalg.qr(rng.standard_normal((an, am)))
    A = U @ np.diag(np.logspace(0, decades, am)) @ V.T
    b = A @ np.abs(rng.standard_normal(an))
    z = rng.standard_normal(an + am) * 10.0 ** decades
    return A, b, [("NonNeg", an), ("Zero", am)], z


def feasibility_problem(am=50, an=100, seed=2):
    """test/testfeasibility.jl shape: find x >= 0 with A x = b (b = A*xsol, xsol >= 0 here so that it
    is feasible), posed on [x; z] as S1 = AffinePlusLinear(A, b, 0, 1) and
    S2 = NonNeg(an) x Zero(am)  (x >= 0, z = 0  <=>  A x = b)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((am, an))
    xsol = np.abs(rng.standard_normal(an))
    b = A @ xsol
    return A, b, [("NonNeg", an), ("Zero", am)]
