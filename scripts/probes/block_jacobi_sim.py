"""CPU simulation (NumPy) of the large-PSD kernel's block one-sided Jacobi: how many outer sweeps does each inner
strategy need?  Same tournament over blocks as k5_psd_hestenes (rr_pair_l), rotations done on the 2*BS x 2*BS Gram
block S = X'X (two-sided on S, accumulated in J, then X <- X J): in exact arithmetic the same rotations as the
column version.  Strategies: cross (today: cross pairs once per step, in-block pairs at step 0), sweepN (N cyclic
sweeps over all pairs of the 2*BS columns per step), full (S diagonalised exactly per step)."""
import sys
import numpy as np


def rr_pair(s, k, D):
    M = D - 1
    if k == 0:
        a, b = D - 1, s
    else:
        a, b = (s + k) % M, (s - k + M) % M
    return (a, b) if a < b else (b, a)


def rot(S, J, p, q, stats):
    al, be, g = S[p, p], S[q, q], S[p, q]
    ab = al * be
    if not ab > 0:
        return
    cos2 = g * g / ab
    stats[0] = max(stats[0], cos2)
    if not g * g > 1e-30 * ab:
        return
    dl = be - al
    hyp = np.sqrt(dl * dl + 4 * g * g)
    t = (2.0 if dl >= 0 else -2.0) * g / (abs(dl) + hyp)
    c = 1 / np.sqrt(1 + t * t)
    s = c * t
    # columns: p' = c p - s q, q' = s p + c q
    for Mx in (S, J):
        cp, cq = Mx[:, p].copy(), Mx[:, q].copy()
        Mx[:, p] = c * cp - s * cq
        Mx[:, q] = s * cp + c * cq
    rp, rq = S[p, :].copy(), S[q, :].copy()
    S[p, :] = c * rp - s * rq
    S[q, :] = s * rp + c * rq


def run(d, BS, strategy, seed=0, warm=None):
    rng = np.random.default_rng(seed)
    G0 = rng.standard_normal((d, d))
    M = (G0 + G0.T) / 2
    sigma = np.linalg.norm(M) * (1 + 1 / 64)
    X = M + sigma * np.eye(d)
    NB = d // BS
    n = 2 * BS
    prev = 1.0
    for sweep in range(40):
        stats = [0.0]
        for step in range(NB - 1):
            for cta in range(NB // 2):
                a, b = rr_pair(step, cta, NB)
                idx = np.r_[a * BS:(a + 1) * BS, b * BS:(b + 1) * BS]
                Xb = X[:, idx]
                S = Xb.T @ Xb
                if strategy == "full":
                    dg = np.sqrt(np.diag(S))
                    C = np.abs(S / np.outer(dg, dg))
                    np.fill_diagonal(C, 0)
                    stats[0] = max(stats[0], C.max() ** 2)
                    w, J = np.linalg.eigh(S)
                else:
                    J = np.eye(n)
                    if strategy == "cross":
                        if step == 0:
                            for st in range(BS - 1):
                                for blk in range(2):
                                    for k in range(BS // 2):
                                        p, q = rr_pair(st, k, BS)
                                        rot(S, J, blk * BS + p, blk * BS + q, stats)
                        for r in range(BS):
                            for w_ in range(BS):
                                rot(S, J, w_, BS + ((w_ + r) % BS), stats)
                    else:
                        ns = int(strategy[5:])
                        for _ in range(ns):
                            for st in range(n - 1):
                                for k in range(n // 2):
                                    p, q = rr_pair(st, k, n)
                                    rot(S, J, p, q, stats)
                X[:, idx] = Xb @ J
        mx = np.sqrt(stats[0])
        print(f"  d={d} BS={BS} {strategy}: sweep {sweep + 1} maxcos {mx:.2e}", flush=True)
        if mx <= 1e-13 or (sweep > 0 and mx <= 3e-8 and mx <= 0.01 * prev):
            return sweep + 1
        prev = mx
    return -1


if __name__ == "__main__":
    d = int(sys.argv[1])
    BS = int(sys.argv[2])
    for strat in sys.argv[3:]:
        print(strat, run(d, BS, strat))
