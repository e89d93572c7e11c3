// batch.cu -- batch mode: one persistent CTA solves one conic problem start to finish (see batch.cuh).
//
// Kernel anatomy (k_batch_solve):
//   * warp CW (the last one) is the producer: an endless cyclic stream of the current problem's A,
//     tile after tile (8 rows x lda, one contiguous cp.async.bulk each), through an S-stage mbarrier
//     ring.  Every pass over A -- CG iteration, initial residual, residual check -- consumes the same
//     tiles in the same order, so the producer never needs to know what the consumers compute and the
//     ring stays full across the vector phases between two passes.
//   * warps 0..CW-1 are consumers: the fused dual mat-vec (A*[x1 x2] and A'*[y1 y2] from one pass,
//     column sums in registers for the whole pass, row sums reduced per tile), then every vector phase
//     of the iteration (KKT epilogue + dots, CG updates, relaxation, cone projection, residuals) as
//     block-strided loops over the problem's vectors (global memory, L1/L2 resident) with
//     deterministic block reductions.  All CG / algorithm scalars are thread-uniform registers.
//   * problems are handed out by a global work counter; on a switch the consumers publish the next
//     problem and the tile sequence number at which it starts, then discard the S speculatively
//     prefetched tiles of the finished problem.
// Element-wise arithmetic is the single-problem path's (kernels.cuh), operation for operation.
#include <algorithm>
#include <cmath>

#include "batch.cuh"
#include "solver.cuh"

namespace fos {

namespace {

struct BK {
    const BatchArgs *a;
    int ct, NT, warp, lane;
    double *tiles, *s_ax, *s_atw, *s_part, *s_wv, *s_red;
    SocScale *s_soc;
    uint64_t *full, *empty;
    uint32_t t;  // tiles consumed so far by this CTA
    int stage;   // ring stage of tile t
    uint32_t phase;
    // current problem
    const double *b, *c;
    const int32_t *drow, *srow_ptr, *srow_id, *scol, *ccol_ptr, *ccol_id, *crow;
    const double *sval, *cval;
    int ns, ncs;
    double *v[BV_COUNT];
    double nb, ncn;
    double *recs;
    // thread-uniform problem state
    int64_t s1_calls;
    int firstrun, cgiter, status, checked, warn, nrec;
    int64_t total_cg, total_passes;
    double alpha12, fista_t;
};

__device__ __forceinline__ void cbar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

template <int NQ>
__device__ __forceinline__ void bt_reduce(BK &k, double (&q)[NQ])
{
#pragma unroll
    for (int i = 0; i < NQ; i++) {
        const double s = warp_sum(q[i]);
        if (k.lane == 0) k.s_red[i * 16 + k.warp] = s;
    }
    cbar(k.NT);
    const int CW = k.NT >> 5;
#pragma unroll
    for (int i = 0; i < NQ; i++) {
        double s = 0.0;
        for (int w = 0; w < CW; w++) s += k.s_red[i * 16 + w];
        q[i] = s;
    }
    cbar(k.NT);
}

// One pass over A: s_ax[v] = A * X_v, s_atw[v] = A' * W_v, v = 0, 1.
// Every tile element is read from shared memory ONCE (round 1 read it twice -- once by a column owner, once by a
// row owner -- plus two X loads per element: the shared-memory pipe sat at 70 % of its peak and bounded the pass;
// profiles/r1_ncu_other_kernels.md).  Thread ct owns the column pairs ct + kk*NT for the whole pass:
//   * the X entries of its columns live in registers (no X staging in shared memory at all);
//   * the four column sums of A'W per pair stay in registers for the whole pass -- no reduction;
//   * its contribution to the 8 x 2 row sums of a tile is reduced over the warp with the transposed butterfly of
//     the single-problem kernel (2*BT_TR values), then over the consumer warps through a double-buffered
//     shared array: one consumer barrier per tile, fixed order -> deterministic.
template <int KP>
__device__ __forceinline__ void bt_pass(BK &k, const double *X0, const double *X1, const double *W0, const double *W1)
{
    const BatchArgs &a = *k.a;
    const int64_t lda = a.lda;
    const int npairs = (int)(lda >> 1);
    const int S = a.S;
    const int64_t tile_elems = (int64_t)BT_TR * lda;
    const int64_t n_pad = a.L.n_pad, m_pad = a.L.m_pad;
    const int mr = a.mr;
    const int CW = k.NT >> 5;
    constexpr int V = 2 * BT_TR;  // row sums per tile
    double2 xv[KP][2];
#pragma unroll
    for (int kk = 0; kk < KP; kk++) {
        const int cp = k.ct + kk * k.NT;
        const bool in = cp < npairs && 2 * (int64_t)cp < n_pad;
        xv[kk][0] = in ? *reinterpret_cast<const double2 *>(X0 + 2 * cp) : make_double2(0.0, 0.0);
        xv[kk][1] = in ? *reinterpret_cast<const double2 *>(X1 + 2 * cp) : make_double2(0.0, 0.0);
    }
    // W rides in shared memory IN TILE-ROW ORDER (gathered through drow once per pass): the inner loop then reads it
    // at compile-time offsets instead of chasing drow -> address per row (a dependent global load per row, 8.5 % of
    // the kernel's stall samples: profiles/r2_c5_notes.md); padding rows of the last tile read 0
    const int nrows_t = a.ntiles * BT_TR;
    for (int t = k.ct; t < nrows_t; t += k.NT) {
        const int rr = __ldg(k.drow + t);
        const bool real = rr < m_pad;
        k.s_wv[t] = real ? W0[rr] : 0.0;
        k.s_wv[mr + t] = real ? W1[rr] : 0.0;
    }
    double2 ca[KP][2];
#pragma unroll
    for (int kk = 0; kk < KP; kk++) ca[kk][0] = ca[kk][1] = make_double2(0.0, 0.0);
    cbar(k.NT);
    for (int tile = 0; tile < a.ntiles; tile++) {
        mbar_wait(&k.full[k.stage], k.phase);
        const double *tp = k.tiles + (size_t)k.stage * tile_elems;
        const int row0 = tile * BT_TR;
        double rs[V];
#pragma unroll
        for (int r = 0; r < BT_TR; r++) {
            const double w0 = k.s_wv[row0 + r], w1 = k.s_wv[mr + row0 + r];
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < KP; kk++) {
                const int cp = k.ct + kk * k.NT;
                if (cp < npairs) {
                    const double2 e = *reinterpret_cast<const double2 *>(tp + (size_t)r * lda + 2 * cp);
                    ca[kk][0].x = fma(e.x, w0, ca[kk][0].x);
                    ca[kk][0].y = fma(e.y, w0, ca[kk][0].y);
                    ca[kk][1].x = fma(e.x, w1, ca[kk][1].x);
                    ca[kk][1].y = fma(e.y, w1, ca[kk][1].y);
                    s0 = fma(e.x, xv[kk][0].x, s0);
                    s0 = fma(e.y, xv[kk][0].y, s0);
                    s1 = fma(e.x, xv[kk][1].x, s1);
                    s1 = fma(e.y, xv[kk][1].y, s1);
                }
            }
            rs[r] = s0;
            rs[BT_TR + r] = s1;
        }
        __syncwarp();
        if (k.lane == 0) mbar_arrive(&k.empty[k.stage]);  // the tile is in registers: hand the stage back early
        warp_transpose_reduce<V>(rs, k.lane);
        // The warps' partial row sums wait in shared memory; the consumers meet once per BT_GROUP tiles (the per-tile
        // barrier was 17 % of the kernel's stall samples, profiles/r2_c5_notes.md).  2 * BT_GROUP buffers: a buffer is
        // written again two groups later, and the barrier of the group in between separates that from its readers.
        double *part = k.s_part + (tile & (2 * BT_GROUP - 1)) * (CW * V);
        constexpr int LPV = 32 / V;  // lanes that end up holding the same value
        if ((k.lane & (LPV - 1)) == 0) part[k.warp * V + (k.lane / LPV)] = rs[0];
        if ((tile & (BT_GROUP - 1)) == BT_GROUP - 1 || tile == a.ntiles - 1) {
            cbar(k.NT);
            const int t0 = tile & ~(BT_GROUP - 1), nt = tile - t0 + 1;
            if (k.ct < V * nt) {
                const int tt = k.ct / V, idx = k.ct - tt * V;
                const double *pt = k.s_part + ((t0 + tt) & (2 * BT_GROUP - 1)) * (CW * V);
                double sum = 0.0;
                for (int w = 0; w < CW; w++) sum += pt[w * V + idx];  // same order as ever: bitwise the per-tile result
                const int v = idx / BT_TR, r = idx - v * BT_TR;
                k.s_ax[v * mr + __ldg(k.drow + (t0 + tt) * BT_TR + r)] = sum;
            }
        }
        k.t++;
        if (++k.stage == S) {
            k.stage = 0;
            k.phase ^= 1;
        }
    }
#pragma unroll
    for (int kk = 0; kk < KP; kk++) {
        const int cp = k.ct + kk * k.NT;
        if (cp < npairs) {
            *reinterpret_cast<double2 *>(k.s_atw + 2 * cp) = ca[kk][0];
            *reinterpret_cast<double2 *>(k.s_atw + n_pad + 2 * cp) = ca[kk][1];
        }
    }
    // sparse rows (CSR): A x
    for (int sr = k.ct; sr < k.ns; sr += k.NT) {
        double a0 = 0.0, a1 = 0.0;
        for (int p = __ldg(k.srow_ptr + sr); p < __ldg(k.srow_ptr + sr + 1); p++) {
            const int j = __ldg(k.scol + p);
            const double v = __ldg(k.sval + p);
            a0 = fma(v, X0[j], a0);
            a1 = fma(v, X1[j], a1);
        }
        const int row = __ldg(k.srow_id + sr);
        k.s_ax[row] = a0;
        k.s_ax[mr + row] = a1;
    }
    cbar(k.NT);
    // columns with sparse entries (CSC): add A' w of the sparse rows to the dense column sums
    for (int sc = k.ct; sc < k.ncs; sc += k.NT) {
        double a0 = 0.0, a1 = 0.0;
        for (int p = __ldg(k.ccol_ptr + sc); p < __ldg(k.ccol_ptr + sc + 1); p++) {
            const int i = __ldg(k.crow + p);
            const double v = __ldg(k.cval + p);
            a0 = fma(v, W0[i], a0);  // W by original row: straight from the (L1-resident) vector
            a1 = fma(v, W1[i], a1);
        }
        const int col = __ldg(k.ccol_id + sc);
        k.s_atw[col] += a0;
        k.s_atw[n_pad + col] += a1;
    }
    k.total_passes++;
    cbar(k.NT);
}

template <int KP>
__device__ __forceinline__ void bt_kkt_pass(BK &k, const double *v)
{
    const Lay &L = k.a->L;
    bt_pass<KP>(k, v, v + L.LP, v + L.n_pad, v + L.LP + L.n_pad);
}

// out = [I Q'; Q -I] in from the pass results (k2_kkt_hsde of kernels.cuh).
//   MODE K2_AP:    out = Ap, returns <Ap, in>
//   MODE K2_RESID: r = rhs - out, p = r, returns <r, r>
template <int MODE>
__device__ __forceinline__ double bt_kkt_epilogue(BK &k, const double *in, double *out,
                                                  const double *rhs, double *r,
                                                  double *p)
{
    const Lay &L = k.a->L;
    const int64_t LP = L.LP, oy = L.n_pad, ot = L.n_pad + L.m_pad;
    const double tau1 = in[ot], tau2 = in[LP + ot];
    const double *ax0 = k.s_ax, *ax1 = k.s_ax + k.a->mr, *atw0 = k.s_atw, *atw1 = k.s_atw + L.n_pad;
    double q[5] = {0, 0, 0, 0, 0};
    for (int64_t base = k.ct; base < ot; base += 4 * (int64_t)k.NT) {
        // loads of four entries first (see the CG update loop), then the arithmetic and the stores
        double i1s[4], i2s[4], cbs[4], r1s[4], r2s[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t e = base + (int64_t)u * k.NT;
            const bool inr = e < ot;
            i1s[u] = inr ? in[e] : 0.0;
            i2s[u] = inr ? in[LP + e] : 0.0;
            cbs[u] = 0.0;
            if (inr) {
                if (e < oy) cbs[u] = e < L.n ? k.c[e] : 0.0;
                else cbs[u] = (e - oy) < L.m ? k.b[e - oy] : 0.0;
            }
            if (MODE == K2_RESID) {
                r1s[u] = inr ? rhs[e] : 0.0;
                r2s[u] = inr ? rhs[LP + e] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t e = base + (int64_t)u * k.NT;
            if (e >= ot) continue;
            double o1 = 0.0, o2 = 0.0;
            const double i1 = i1s[u], i2 = i2s[u];
            if (e < oy) {
                if (e < L.n) {
                    const double cj = cbs[u];
                    const double q1 = add_(atw0[e], mul_(tau1, cj));  // HSDEAffine.jl:51,54
                    const double q2 = add_(atw1[e], mul_(tau2, cj));
                    o1 = add_(-q2, i1);  // Q'in2 + in1 (HSDEAffine.jl:61-65)
                    o2 = sub_(q1, i2);   // Q in1 - in2
                    q[0] = fma(cj, i1, q[0]);
                    q[2] = fma(cj, i2, q[2]);
                }
            } else {
                const int64_t i = e - oy;
                if (i < L.m) {
                    const double bi = cbs[u];
                    const double q1 = -sub_(ax0[i], mul_(tau1, bi));  // HSDEAffine.jl:52,55,56
                    const double q2 = -sub_(ax1[i], mul_(tau2, bi));
                    o1 = add_(-q2, i1);
                    o2 = sub_(q1, i2);
                    q[1] = fma(bi, i1, q[1]);
                    q[3] = fma(bi, i2, q[3]);
                }
            }
            if (MODE == K2_AP) {
                out[e] = o1;
                out[LP + e] = o2;
                q[4] = fma(o1, i1, q[4]);
                q[4] = fma(o2, i2, q[4]);
            } else {
                const double r1 = sub_(r1s[u], o1), r2 = sub_(r2s[u], o2);
                r[e] = r1;
                r[LP + e] = r2;
                p[e] = r1;
                p[LP + e] = r2;
                q[4] = fma(r1, r1, q[4]);
                q[4] = fma(r2, r2, q[4]);
            }
        }
    }
    bt_reduce<5>(k, q);
    const double q1t = sub_(-q[0], q[1]);  // HSDEAffine.jl:57
    const double q2t = sub_(-q[2], q[3]);
    const double o1 = add_(-q2t, tau1);
    const double o2 = sub_(q1t, tau2);
    double res;
    if (MODE == K2_AP) {
        if (k.ct == 0) {
            out[ot] = o1;
            out[LP + ot] = o2;
        }
        res = q[4] + o1 * tau1 + o2 * tau2;
    } else {
        const double r1 = sub_(rhs[ot], o1), r2 = sub_(rhs[LP + ot], o2);
        if (k.ct == 0) {
            r[ot] = r1;
            r[LP + ot] = r2;
            p[ot] = r1;
            p[LP + ot] = r2;
        }
        res = q[4] + r1 * r1 + r2 * r2;
    }
    cbar(k.NT);
    return res;
}

// prox!(y, S::AffinePlusLinear, x) (affinepluslinear.jl:83-126) with the fused right-hand side
// (k_fuse_prep): result in BV_SOL.  CG = conjugategradients.jl:31-55, max_iters = 1000.
template <int KP>
__device__ __forceinline__ void bt_s1_prox(BK &k, const double *xin)
{
    const BatchArgs &a = *k.a;
    const Lay &L = a.L;
    const int64_t NP = L.NP, LP = L.LP;
    double *sol = k.v[BV_SOL], *r = k.v[BV_R], *p = k.v[BV_P], *Ap = k.v[BV_AP];
    if (k.firstrun) {  // :101-104
        for (int64_t e = k.ct; e < NP; e += k.NT) sol[e] = xin[e];
        k.firstrun = 0;
        cbar(k.NT);
    }
    // d = [x0_1 ; x0_2 - x2] into Ap (k_fuse_prep; HSDE: beta = 1, q = b = 0, so rhs' = xin)
    for (int64_t e = k.ct; e < NP; e += k.NT) Ap[e] = e < LP ? sol[e] : sub_(sol[e], mul_(1.0, xin[e]));
    cbar(k.NT);
    const double tol = k.s1_calls <= a.tol_n ? a.tol_table[k.s1_calls - 1] : a.tol_floor;  // :108-112
    k.s1_calls += 1;                                                                        // :114
    const int max_iters = 1000;
    bt_kkt_pass<KP>(k, Ap);
    double rn = bt_kkt_epilogue<K2_RESID>(k, Ap, nullptr, xin, r, p);  // cg :32-36
    int iter = 1;
    for (;;) {
        bt_kkt_pass<KP>(k, p);
        const double pAp = bt_kkt_epilogue<K2_AP>(k, p, Ap, nullptr, nullptr, nullptr);
        const double alpha = rn / pAp;  // :39
        double q[1] = {0.0};
        // four entries per thread at a time, all loads first: the stores of one entry would otherwise fence the
        // loads of the next (possible aliasing) and every entry would pay its own L2 round trip
        for (int64_t base = k.ct; base < NP; base += 4 * (int64_t)k.NT) {
            double ps[4], ss[4], rs[4], as[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t e = base + (int64_t)u * k.NT;
                const bool in = e < NP;
                ps[u] = in ? p[e] : 0.0;
                ss[u] = in ? sol[e] : 0.0;
                rs[u] = in ? r[e] : 0.0;
                as[u] = in ? Ap[e] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t e = base + (int64_t)u * k.NT;
                if (e < NP) {
                    sol[e] = add_(ss[u], mul_(alpha, ps[u]));          // :40
                    const double re = sub_(rs[u], mul_(alpha, as[u]));  // :41
                    r[e] = re;
                    q[0] = fma(re, re, q[0]);
                }
            }
        }
        bt_reduce<1>(k, q);
        const double rr = q[0];
        const double rnorm = sqrt(rr);
        if (rnorm <= tol || iter >= max_iters) {  // :42
            if (iter >= max_iters) k.warn = 1;    // :53
            break;
        }
        const double beta = rr / rn;  // :45-47
        rn = rr;
        iter += 1;  // :51
        for (int64_t base = k.ct; base < NP; base += 4 * (int64_t)k.NT) {  // p = beta p + r  (:49-50)
            double ps[4], rs[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t e = base + (int64_t)u * k.NT;
                ps[u] = e < NP ? p[e] : 0.0;
                rs[u] = e < NP ? r[e] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t e = base + (int64_t)u * k.NT;
                if (e < NP) p[e] = add_(mul_(beta, ps[u]), rs[u]);
            }
        }
        cbar(k.NT);
    }
    k.cgiter = iter;
    k.total_cg += iter;
    cbar(k.NT);
}

// squared tail norms of the SOC cones -> classification (k4_soc_norms)
__device__ __forceinline__ void bt_soc_norms(BK &k, const double *in)
{
    const BatchArgs &a = *k.a;
    for (int ci = 0; ci < a.nsoc; ci++) {
        const SocCone K = a.soc[ci];
        double q[1] = {0.0};
        for (int64_t j = k.ct; j < K.len - 1; j += k.NT) {
            const double w = in[K.head + 1 + j];
            q[0] = fma(w, w, q[0]);
        }
        bt_reduce<1>(k, q);
        if (k.ct == 0) {
            const double nx = sqrt(q[0]);
            double t = in[K.head];
            if (K.dual) t = -t;
            SocScale R;
            R.pad_ = 0;
            R.nx = nx;
            R.rho = 0.0;
            if (t <= -nx) R.mode = 0;
            else if (t >= nx) R.mode = 1;
            else {
                R.mode = 2;
                R.rho = 0.5 * (1.0 + t / nx);
            }
            k.s_soc[ci] = R;
        }
    }
    cbar(k.NT);
}

// proj = P_S2(in) followed by the algorithm's epilogue (k4_cone_apply)
template <int EPI>
__device__ __forceinline__ void bt_cone_apply(BK &k, const double *in, double *proj,
                                              const EpiArgs &E)
{
    const BatchArgs &a = *k.a;
    const int64_t NP = a.L.NP;
    bt_soc_norms(k, in);
    double a2 = E.a2, om_a2 = E.om_a2;
    if (EPI == EPI_GAPA) {
        a2 = k.alpha12;
        om_a2 = 1.0 - a2;
    }
    double q[3] = {0, 0, 0};
    for (int64_t e = k.ct; e < NP; e += k.NT) {
        const double t1 = in[e];
        const double pj = cone_project(a.ops[e], t1, proj, e, a.cone_of, k.s_soc);
        proj[e] = pj;
        if (EPI == EPI_GAP || EPI == EPI_GAPA) {
            const double t2 = add_(mul_(a2, pj), mul_(om_a2, t1));
            E.tmp2[e] = t2;
            const double xo = E.x[e];
            if (EPI == EPI_GAPA) {
                const double d1 = sub_(t2, t1), d2 = sub_(t1, xo);
                q[0] = fma(d1, d2, q[0]);
                q[1] = fma(d1, d1, q[1]);
                q[2] = fma(d2, d2, q[2]);
            }
            E.x[e] = add_(mul_(E.a, t2), mul_(E.om_a, xo));
        } else if (EPI == EPI_FISTA) {
            const double xo = E.x[e];
            E.aux1[e] = xo;
            E.x[e] = pj;
            E.aux2[e] = add_(pj, mul_(E.coef, sub_(pj, xo)));
        } else if (EPI == EPI_DYKSTRA) {
            E.x[e] = pj;
            E.aux1[e] = sub_(t1, pj);
        }
    }
    if (EPI == EPI_GAPA) {
        bt_reduce<3>(k, q);
        double scl = fabs(q[0]) / sqrt(q[1] * q[2]);  // normedScalar, gapa.jl:47
        if (isnan(scl)) scl = 0.0;
        else scl = scl < 0.0 ? 0.0 : (scl > 1.0 ? 1.0 : scl);
        const double s = sqrt(1.0 - scl * scl);
        const double aopt = 2.0 / (1.0 + s);
        k.alpha12 = (1.0 - E.betaA) * aopt + E.betaA * 2.0;
    }
    cbar(k.NT);
}

// checkstatus (HSDEStatus.jl:27-71) on z: one more pass over A (k6_check_hsde)
template <int KP>
__device__ __forceinline__ void bt_check(BK &k, const double *z, int64_t iter_i)
{
    const BatchArgs &a = *k.a;
    const Lay &L = a.L;
    const int64_t LP = L.LP, oy = L.n_pad, ot = L.n_pad + L.m_pad;
    bt_pass<KP>(k, z, z, z + L.n_pad, z + L.n_pad);
    const double tau = z[ot], kap = z[LP + ot];
    double q[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t e = k.ct; e < ot; e += k.NT) {
        if (e < oy) {
            if (e < L.n) {
                const double aty = k.s_atw[e];
                const double x = z[e], r = z[LP + e], cj = k.c[e];
                const double dv = sub_(add_(aty / tau, cj), r / tau);  // :35
                q[1] = fma(dv, dv, q[1]);
                q[2] = fma(cj, x, q[2]);
                q[5] = fma(aty, aty, q[5]);
            }
        } else {
            const int64_t i = e - oy;
            if (i < L.m) {
                const double ax = k.s_ax[i];
                const double y = z[e], s = z[LP + e], bi = k.b[i];
                const double pv = sub_(add_(ax / tau, s / tau), bi);  // :34
                q[0] = fma(pv, pv, q[0]);
                q[3] = fma(bi, y, q[3]);
                const double uv = add_(ax, s);
                q[4] = fma(uv, uv, q[4]);
            }
        }
    }
    bt_reduce<6>(k, q);
    const double nb = k.nb, ncn = k.ncn, eps = a.eps;
    const double p = sqrt(q[0]) / fabs(1.0 + nb);
    const double d = sqrt(q[1]) / fabs(1.0 + ncn);
    const double ctx = q[2], bty = q[3];
    const double g = fabs(ctx / tau + bty / tau) / (1.0 + fabs(ctx / tau) + fabs(bty / tau));  // :38
    int status = FOS_STATUS_CONTINUE;
    if (p <= eps * (1.0 + nb) && d <= eps * (1.0 + ncn) && g <= eps * (1.0 + fabs(ctx / tau) + fabs(bty / tau)))
        status = FOS_STATUS_OPTIMAL;  // :54
    else if (sqrt(q[4]) <= eps * (-ctx / ncn))
        status = FOS_STATUS_UNBOUNDED;  // :59
    else if (sqrt(q[5]) <= eps * (-bty / nb))
        status = FOS_STATUS_INFEASIBLE;  // :61
    if (k.ct == 0 && k.nrec < a.rec_cap) {
        double *R = k.recs + (size_t)k.nrec * FOS_REC_LEN;
        R[0] = (double)iter_i; R[1] = p; R[2] = d; R[3] = g; R[4] = ctx; R[5] = bty; R[6] = kap; R[7] = tau;
        R[8] = (double)k.cgiter;
        R[9] = (double)status;
    }
    k.nrec += 1;
    k.status = status;
    k.checked = 1;
    cbar(k.NT);
}

// one iteration of the selected algorithm (Handle::step of solver.cu; GAPP is not offered in batch mode)
template <int KP>
__device__ __forceinline__ void bt_step(BK &k, int64_t i)
{
    const BatchArgs &a = *k.a;
    const int64_t NP = a.L.NP;
    double *x = k.v[BV_X], *sol = k.v[BV_SOL], *tmp1 = k.v[BV_TMP1], *proj = k.v[BV_PROJ];
    EpiArgs E{};
    E.tmp2 = k.v[BV_TMP2];
    E.x = x;
    E.betaA = a.betaA;
    const bool due = (i % a.checki) == 0;
    switch (a.alg) {
    case FOS_ALG_GAP:
    case FOS_ALG_GAPA: {
        const bool ada = a.alg == FOS_ALG_GAPA;
        bt_s1_prox<KP>(k, x);  // gap.jl:45
        const double a1 = ada ? k.alpha12 : a.alpha1;
        const double b1 = 1.0 - a1;
        for (int64_t e = k.ct; e < NP; e += k.NT) tmp1[e] = add_(mul_(a1, sol[e]), mul_(b1, x[e]));  // :48
        cbar(k.NT);
        E.a2 = a.alpha2;
        E.om_a2 = 1.0 - a.alpha2;
        E.a = a.alpha;
        E.om_a = 1.0 - a.alpha;
        if (ada) bt_cone_apply<EPI_GAPA>(k, tmp1, proj, E);
        else bt_cone_apply<EPI_GAP>(k, tmp1, proj, E);
        break;
    }
    case FOS_ALG_FISTA: {
        double *fy = k.v[BV_FY];
        if (i == 1) {  // fista.jl:31-33
            for (int64_t e = k.ct; e < NP; e += k.NT) fy[e] = x[e];
            cbar(k.NT);
        }
        bt_s1_prox<KP>(k, fy);  // :35
        const double al = a.alpha, bl = 1.0 - a.alpha;
        for (int64_t e = k.ct; e < NP; e += k.NT) tmp1[e] = add_(mul_(al, sol[e]), mul_(bl, fy[e]));  // :37
        cbar(k.NT);
        const double told = k.fista_t;
        k.fista_t = __ddiv_rn(__dadd_rn(1.0, sqrt(__dadd_rn(1.0, __dmul_rn(__dmul_rn(4.0, told), told)))), 2.0);  // :45
        E.coef = __ddiv_rn(__dsub_rn(told, 1.0), k.fista_t);  // :46
        E.aux1 = k.v[BV_FXOLD];
        E.aux2 = fy;
        bt_cone_apply<EPI_FISTA>(k, tmp1, proj, E);  // :39-41,46
        break;
    }
    case FOS_ALG_DYKSTRA: {
        double *w1 = k.v[BV_W1], *w2 = k.v[BV_W2], *dp = k.v[BV_DP], *dq = k.v[BV_DQ], *dy = k.v[BV_DY];
        for (int64_t e = k.ct; e < NP; e += k.NT) w1[e] = add_(x[e], mul_(1.0, dp[e]));  // x + p
        cbar(k.NT);
        bt_s1_prox<KP>(k, w1);  // dykstra.jl:29
        for (int64_t e = k.ct; e < NP; e += k.NT) {
            const double ye = sol[e];
            dy[e] = ye;
            dp[e] = sub_(w1[e], ye);  // :31
            w2[e] = add_(ye, dq[e]);  // y + q
        }
        cbar(k.NT);
        E.aux1 = dq;
        bt_cone_apply<EPI_DYKSTRA>(k, w2, proj, E);  // :32,35
        break;
    }
    default: break;
    }
    if (due) bt_check<KP>(k, proj, i);  // on the unrelaxed S2 projection (gap.jl:56)
    else k.checked = 0;                 // HSDEStatus.jl:66-69
}

template <int KP, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_batch_solve(const BatchArgs a)
{
    extern __shared__ __align__(128) unsigned char bt_smem[];
    const int64_t tile_elems = (int64_t)BT_TR * a.lda;
    double *tiles = reinterpret_cast<double *>(bt_smem);
    double *s_ax = tiles + (size_t)a.S * tile_elems;
    double *s_atw = s_ax + 2 * a.mr;
    double *s_wv = s_atw + 2 * a.L.n_pad;
    double *s_part = s_wv + 2 * a.mr;           // [2 * BT_GROUP tiles][CW warps][2 * BT_TR] row-sum partials
    double *s_red = s_part + 2 * BT_GROUP * a.CW * 2 * BT_TR;
    SocScale *s_soc = reinterpret_cast<SocScale *>(s_red + 8 * 16);
    uint64_t *full = reinterpret_cast<uint64_t *>(s_soc + BT_MAX_SOC);
    uint64_t *empty = full + BT_MAX_STAGES;
    volatile int *s_next_pb = reinterpret_cast<volatile int *>(empty + BT_MAX_STAGES);
    volatile uint32_t *s_switch_at = reinterpret_cast<volatile uint32_t *>(s_next_pb + 1);

    const int CW = a.CW, NT = CW * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CW);
        }
        *s_next_pb = -1;
        *s_switch_at = 0xffffffffu;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int64_t e = threadIdx.x; e < 4 * (int64_t)a.mr + 2 * a.L.n_pad; e += blockDim.x) s_ax[e] = 0.0;  // ax, atw, wv
    __syncthreads();

    if (warp == CW) {
        // ================= producer: cyclic stream of the current problem's tiles =================
        if (lane == 0) {
            const uint32_t tile_bytes = (uint32_t)(tile_elems * sizeof(double));
            const double *base = nullptr;
            int tix = 0;
            int s = 0;
            uint32_t eph = 0;  // parity to wait for on empty[s] (first used by tiles S..2S-1); flips every S tiles
            const uint64_t stream_once = l2_policy_evict_first();  // the active problems' tiles (296 x 1 MB) never fit L2
            for (uint32_t T = 0;; T++) {
                if (T >= (uint32_t)a.S) mbar_wait(&empty[s], eph);
                while (base == nullptr && *s_switch_at == 0xffffffffu) {
                }  // the very first problem has not been published yet
                if (T == *s_switch_at) {
                    const int pb = *s_next_pb;
                    if (pb < 0) break;
                    base = a.A + (size_t)pb * a.a_stride;
                    tix = 0;
                }
                mbar_expect_tx(&full[s], tile_bytes);
                bulk_load_1d_hint(tiles + (size_t)s * tile_elems, base + (size_t)tix * tile_elems, tile_bytes, &full[s], stream_once);
                tix = tix + 1 == a.ntiles ? 0 : tix + 1;
                if (++s == a.S) {
                    s = 0;
                    if (T >= (uint32_t)a.S) eph ^= 1;
                }
            }
        }
        return;
    }

    // ================= consumers =================
    __shared__ int s_pb;
    BK k;
    k.a = &a;
    k.ct = threadIdx.x;
    k.NT = NT;
    k.warp = warp;
    k.lane = lane;
    k.tiles = tiles;
    k.s_ax = s_ax;
    k.s_atw = s_atw;
    k.s_part = s_part;
    k.s_wv = s_wv;
    k.s_red = s_red;
    k.s_soc = s_soc;
    k.full = full;
    k.empty = empty;
    k.t = 0;
    k.stage = 0;
    k.phase = 0;
    bool streaming = false;  // the producer has (or will have) tiles of some problem in flight
    for (;;) {
        if (k.ct == 0) {
            const unsigned int nxt = atomicAdd(a.work_counter, 1u);
            const int pb = nxt < (unsigned int)a.B ? (int)nxt : -1;
            s_pb = pb;
            *s_next_pb = pb;
            __threadfence_block();
            *s_switch_at = streaming ? k.t + (uint32_t)a.S : 0u;
        }
        cbar(NT);
        const int pb = s_pb;
        if (streaming) {
            // discard the S speculatively prefetched tiles of the finished problem
            for (int d = 0; d < a.S; d++) {
                mbar_wait(&full[k.stage], k.phase);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[k.stage]);
                k.t++;
                if (++k.stage == a.S) {
                    k.stage = 0;
                    k.phase ^= 1;
                }
            }
        }
        if (pb < 0) break;
        streaming = true;

        // ---- bind the problem ----
        const Lay &L = a.L;
        k.b = a.b + (size_t)pb * L.m_pad;
        k.c = a.c + (size_t)pb * L.n_pad;
        k.drow = a.drow + (size_t)pb * a.ntiles * BT_TR;
        k.ns = a.sp_count[2 * pb];
        k.ncs = a.sp_count[2 * pb + 1];
        k.srow_ptr = a.srow_ptr + (size_t)pb * (a.NS + 1);
        k.srow_id = a.srow_id + (size_t)pb * a.NS;
        k.scol = a.scol + (size_t)pb * a.NZ;
        k.sval = a.sval + (size_t)pb * a.NZ;
        k.ccol_ptr = a.ccol_ptr + (size_t)pb * (a.NCS + 1);
        k.ccol_id = a.ccol_id + (size_t)pb * a.NCS;
        k.crow = a.crow + (size_t)pb * a.NZ;
        k.cval = a.cval + (size_t)pb * a.NZ;
        k.nb = a.nb[pb];
        k.ncn = a.ncn[pb];
        for (int j = 0; j < BV_COUNT; j++) k.v[j] = a.vec + ((size_t)pb * BV_COUNT + j) * L.NP;
        k.recs = a.recs + (size_t)pb * a.rec_cap * FOS_REC_LEN;
        BatchCtl *ctl = a.ctl + pb;
        k.s1_calls = ctl->s1_calls;
        k.firstrun = ctl->firstrun;
        k.cgiter = ctl->cgiter;
        k.status = ctl->status;
        k.checked = ctl->checked;
        k.warn = ctl->warn_maxit;
        k.alpha12 = ctl->alpha12;
        k.fista_t = ctl->fista_t;
        k.total_cg = ctl->total_cg;
        k.total_passes = ctl->total_passes;
        k.nrec = 0;
        int64_t last_i = ctl->last_i, done = 0;

        if (a.do_run) {
            for (int64_t i = a.i_start; i < a.i_start + a.n_iters && k.status == FOS_STATUS_CONTINUE; i++) {
                last_i = i;           // solverwrapper.jl:24
                bt_step<KP>(k, i);    // :25
                done++;
            }
        }
        if (a.do_finish) {
            // getsol = P2(P1(x)) (gap.jl:82-87 and the identical methods of the other algorithms), then the
            // forced final check when the last iteration was not a check (solverwrapper.jl:31-34)
            bt_s1_prox<KP>(k, k.v[BV_X]);
            double *tmp1 = k.v[BV_TMP1], *sol = k.v[BV_SOL];
            for (int64_t e = k.ct; e < L.NP; e += NT) tmp1[e] = sol[e];
            cbar(NT);
            EpiArgs E{};
            bt_cone_apply<EPI_NONE>(k, tmp1, k.v[BV_PROJ], E);
            if (!k.checked) bt_check<KP>(k, k.v[BV_PROJ], last_i);
        }
        if (k.ct == 0) {
            ctl->s1_calls = k.s1_calls;
            ctl->firstrun = k.firstrun;
            ctl->cgiter = k.cgiter;
            ctl->status = k.status;
            ctl->checked = k.checked;
            ctl->warn_maxit = k.warn;
            ctl->alpha12 = k.alpha12;
            ctl->fista_t = k.fista_t;
            ctl->total_cg = k.total_cg;
            ctl->total_passes = k.total_passes;
            ctl->nrec = k.nrec;
            ctl->last_i = last_i;
            ctl->iters_done = done;
        }
        cbar(NT);
    }
}

// pack / unpack with a batch dimension (blockIdx.y = problem)
__global__ void __launch_bounds__(VBLOCK)
k_pack_batch(SegMap M, const double *logical, int64_t N, double *padded, int64_t NP,
             int64_t pstride)
{
    const double *src = logical + (size_t)blockIdx.y * N;
    double *dst = padded + (size_t)blockIdx.y * pstride;
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        double v = 0.0;
        for (int s = 0; s < M.nseg; s++) {
            const int64_t kk = e - M.dst[s];
            if (kk >= 0 && kk < M.len[s]) v = src[M.src[s] + kk];
        }
        dst[e] = v;
    }
}
__global__ void __launch_bounds__(VBLOCK)
k_unpack_batch(SegMap M, const double *padded, int64_t pstride, double *logical, int64_t N)
{
    const double *src = padded + (size_t)blockIdx.y * pstride;
    double *dst = logical + (size_t)blockIdx.y * N;
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < N; e += (int64_t)gridDim.x * VBLOCK) {
        for (int s = 0; s < M.nseg; s++) {
            const int64_t kk = e - M.src[s];
            if (kk >= 0 && kk < M.len[s]) dst[e] = src[M.dst[s] + kk];
        }
    }
}
// non-zeros per row of every problem: one warp per row
__global__ void __launch_bounds__(256)
k_row_nnz(const double *__restrict__ src, int64_t lda_src, int64_t pstride_src, int64_t m, int64_t n, int32_t *__restrict__ nnz)
{
    const int64_t pb = blockIdx.y;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    const int lane = threadIdx.x & 31;
    const double *r = src + (size_t)pb * pstride_src + (size_t)row * lda_src;
    int c = 0;
    for (int64_t j = lane; j < n; j += 32) c += r[j] != 0.0 ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) nnz[(size_t)pb * m + row] = c;
}
// tile row t of problem pb <- row drow[pb][t] of its matrix (zero row for padding slots)
__global__ void __launch_bounds__(128)
k_gather_dense_rows(const double *__restrict__ src, int64_t lda_src, int64_t pstride_src, int64_t m, int64_t n,
                    const int32_t *__restrict__ drow, int rows_pad, double *__restrict__ dst, int64_t lda, int64_t a_stride)
{
    const int64_t pb = blockIdx.y, t = blockIdx.x;
    const int32_t row = drow[(size_t)pb * rows_pad + t];
    double *d = dst + (size_t)pb * a_stride + (size_t)t * lda;
    const double *s = src + (size_t)pb * pstride_src + (size_t)row * lda_src;
    for (int64_t j = threadIdx.x; j < lda; j += 128) d[j] = (row < m && j < n) ? s[j] : 0.0;
}
// CSR of the sparse rows: one warp per row, non-zeros compacted in column order (ballot prefix)
__global__ void __launch_bounds__(32)
k_extract_sparse_rows(const double *__restrict__ src, int64_t lda_src, int64_t pstride_src, int64_t n,
                      const int32_t *__restrict__ srow_id, const int32_t *__restrict__ srow_ptr, int NS,
                      int32_t *__restrict__ scol, double *__restrict__ sval, int NZ)
{
    const int64_t pb = blockIdx.y;
    const int sr = blockIdx.x;
    const int32_t p0 = srow_ptr[(size_t)pb * (NS + 1) + sr], p1 = srow_ptr[(size_t)pb * (NS + 1) + sr + 1];
    if (p1 == p0) return;  // padding slot
    const int lane = threadIdx.x;
    const double *r = src + (size_t)pb * pstride_src + (size_t)srow_id[(size_t)pb * NS + sr] * lda_src;
    int32_t pos = p0;
    for (int64_t j0 = 0; j0 < n; j0 += 32) {
        const int64_t j = j0 + lane;
        const double v = j < n ? r[j] : 0.0;
        const unsigned int msk = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) {
            const int32_t q = pos + __popc(msk & ((1u << lane) - 1u));
            scol[(size_t)pb * NZ + q] = (int32_t)j;
            sval[(size_t)pb * NZ + q] = v;
        }
        pos += __popc(msk);
    }
}
__global__ void k_batch_reset(BatchCtl *ctl, int B, int what)
{
    // what 0: full reset (load); 1: begin_solve (status object only); 2: set_algorithm (algorithm data only)
    const int pb = blockIdx.x * blockDim.x + threadIdx.x;
    if (pb >= B) return;
    BatchCtl &c = ctl[pb];
    if (what == 0) {
        c.s1_calls = 1;
        c.firstrun = 1;
        c.cgiter = 0;
        c.warn_maxit = 0;
        c.total_cg = 0;
        c.total_passes = 0;
        c.last_i = 0;
        c.iters_done = 0;
        c.nrec = 0;
    }
    if (what == 0 || what == 1) {
        c.status = FOS_STATUS_CONTINUE;
        c.checked = 0;
        c.warn_maxit = 0;  // the warning of conjugategradients.jl:53 is per occurrence
        c.last_i = 0;      // a fresh status object counts from i = 0
    }
    if (what == 0 || what == 2) {
        c.alpha12 = 2.0;  // gapa.jl:29
        c.fista_t = 1.0;  // fista.jl:24
    }
}

}  // namespace

// =======================================================================================
// host side
// =======================================================================================
// Kernel geometry of one problem shape: tile layout of A, ring depth, consumer warps, column pairs per
// consumer thread, CTAs per SM.  Pure host arithmetic (unit-tested on CPU through fos_batch_plan).
BatchGeom batch_geometry(int64_t m, int64_t n)
{
    FOS_REQUIRE(m >= 1 && n >= 1, "empty problem");
    BatchGeom g;
    const int64_t n_pad = ru(n, PAD), m_pad = ru(m, PAD);
    g.lda = n_pad;
    if (g.lda > BT_MAX_LDA)
        throw Error(FOS_ERR_UNSUPPORTED, "batch mode supports n <= " + std::to_string(BT_MAX_LDA) +
                                             " (one CTA per problem); use one handle per problem for larger ones");
    g.ntiles = (int)((m + BT_TR - 1) / BT_TR);
    g.a_stride = (int64_t)g.ntiles * BT_TR * g.lda;
    const int npairs = (int)(g.lda / 2);
    g.KP = 1;
    while (g.KP < BT_MAX_KP && (npairs + g.KP * 32 * BT_MAX_CW - 1) / (g.KP * 32 * BT_MAX_CW) > 1) g.KP++;
    if ((int64_t)g.KP * 32 * BT_MAX_CW < npairs) throw Error(FOS_ERR_UNSUPPORTED, "batch mode: n too large");
    g.CW = std::max(4, (npairs + g.KP * 32 - 1) / (g.KP * 32));
    const size_t tile_bytes = (size_t)BT_TR * g.lda * 8;
    const size_t fixed = (size_t)(4 * (m_pad + 16) + 2 * n_pad + 2 * BT_GROUP * g.CW * 2 * BT_TR + 8 * 16) * 8 +
                         BT_MAX_SOC * sizeof(SocScale) + 2 * BT_MAX_STAGES * 8 + 64;
    // two CTAs (= two problems) per SM when they fit: one CTA's vector phases overlap the other's pass
    g.ctas_per_sm = 1;
    g.S = BT_MAX_STAGES;
    if (g.KP == 1 && (g.CW + 1) * 32 <= 320 && 2 * tile_bytes + fixed <= (size_t)112 * 1024) {
        g.ctas_per_sm = 2;
        while (g.S > 2 && (size_t)g.S * tile_bytes + fixed > (size_t)112 * 1024) g.S--;
    } else {
        while (g.S > 2 && (size_t)g.S * tile_bytes + fixed > (size_t)220 * 1024) g.S--;
    }
    g.smem_bytes = (size_t)g.S * tile_bytes + fixed;
    if (g.smem_bytes > (size_t)227 * 1024)
        throw Error(FOS_ERR_UNSUPPORTED, "batch mode: problem does not fit one SM's shared memory");
    return g;
}

void BatchSolver::load(Handle *h_, int64_t B_, int64_t m, int64_t n, const double *A, int64_t lda_src,
                       int64_t pstride_src, int location, const double *b, const double *c, int64_t nc1,
                       const int32_t *t1, const int64_t *l1, int64_t nc2, const int32_t *t2, const int64_t *l2)
{
    h = h_;
    FOS_REQUIRE(B_ >= 1 && B_ <= (int64_t)1 << 24, "batch size out of range");
    FOS_REQUIRE(m >= 1 && n >= 1, "empty problem");
    B = B_;
    int64_t s1 = 0, s2 = 0;
    for (int64_t k = 0; k < nc1; k++) s1 += l1[k];
    for (int64_t k = 0; k < nc2; k++) s2 += l2[k];
    FOS_REQUIRE(s1 == m, "constraint cones do not cover 1:m (cones.jl:66-72)");
    FOS_REQUIRE(s2 == n, "variable cones do not cover 1:n (cones.jl:66-72)");
    L.form = 0;
    L.n = n;
    L.m = m;
    L.n_pad = ru(n, PAD);
    L.m_pad = ru(m, PAD);
    L.LP = L.n_pad + L.m_pad + PAD;
    L.NP = 2 * L.LP;
    const int64_t l = m + n + 1;
    N = 2 * l;
    seg.nseg = 6;
    const int64_t lens[6] = {n, m, 1, n, m, 1};
    const int64_t srcs[6] = {0, n, n + m, l, l + n, l + n + m};
    const int64_t dsts[6] = {0, L.n_pad, L.n_pad + L.m_pad, L.LP, L.LP + L.n_pad, L.LP + L.n_pad + L.m_pad};
    for (int k = 0; k < 6; k++) {
        seg.len[k] = lens[k];
        seg.src[k] = srcs[k];
        seg.dst[k] = dsts[k];
    }
    const BatchGeom g = batch_geometry(m, n);
    lda = g.lda;
    ntiles = g.ntiles;
    a_stride = g.a_stride;
    KP = g.KP;
    CW = g.CW;
    S = g.S;
    ctas_per_sm = g.ctas_per_sm;
    smem_bytes = g.smem_bytes;

    // matrices: classify the rows of every problem (dense / sparse / empty), gather the dense rows into tiles,
    // extract the sparse rows as CSR on the device and build their CSC on the host (small)
    {
        DevBuf<double> whole;
        const double *src = A;
        if (location == FOS_MEM_HOST) {
            const size_t elems = (size_t)(B - 1) * pstride_src + (size_t)(m - 1) * lda_src + n;
            if (elems * 8 > ((size_t)96 << 30))
                throw Error(FOS_ERR_NOMEM, "host batch larger than 96 GB: hand the matrices over in device memory");
            whole.alloc(elems, false);
            FOS_CUDA(cudaMemcpy(whole.p, A, elems * 8, cudaMemcpyHostToDevice));
            FOS_SYNC_LEGACY();
            src = whole.p;
        }
        DevBuf<int32_t> d_nnz;
        d_nnz.alloc((size_t)B * m);
        k_row_nnz<<<dim3((unsigned)((m + 7) / 8), (unsigned)B), 256, 0, h->stream>>>(src, lda_src, pstride_src, m, n, d_nnz.p);
        std::vector<int32_t> nnz((size_t)B * m);
        // the handle's stream is non-blocking: a synchronous cudaMemcpy (legacy stream) does not wait for it
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        FOS_CUDA(cudaMemcpy(nnz.data(), d_nnz.p, nnz.size() * 4, cudaMemcpyDeviceToHost));
        const int64_t thr = hybrid ? std::max<int64_t>(1, n / 8) : 0;  // rows with <= thr non-zeros go to CSR
        std::vector<std::vector<int32_t>> drows((size_t)B), srows((size_t)B);
        int64_t ND = 0, NSmax = 0, NZmax = 0;
        dense_rows_total = 0;
        sparse_nnz_total = 0;
        for (int64_t p = 0; p < B; p++) {
            int64_t nz = 0;
            for (int64_t i = 0; i < m; i++) {
                const int32_t c = nnz[(size_t)p * m + i];
                if (c == 0 && hybrid) continue;  // an empty row contributes nothing to A x or A' w
                if (hybrid && c <= thr) {
                    srows[(size_t)p].push_back((int32_t)i);
                    nz += c;
                } else {
                    drows[(size_t)p].push_back((int32_t)i);
                }
            }
            ND = std::max<int64_t>(ND, (int64_t)drows[(size_t)p].size());
            NSmax = std::max<int64_t>(NSmax, (int64_t)srows[(size_t)p].size());
            NZmax = std::max<int64_t>(NZmax, nz);
            dense_rows_total += (int64_t)drows[(size_t)p].size();
            sparse_nnz_total += nz;
        }
        ntiles = (int)std::max<int64_t>(1, (ND + BT_TR - 1) / BT_TR);
        a_stride = (int64_t)ntiles * BT_TR * lda;
        NS = (int32_t)std::max<int64_t>(NSmax, 1);
        NZ = (int32_t)std::max<int64_t>(NZmax, 1);
        const int32_t dummy = (int32_t)L.m_pad;  // a row slot whose W entry is always zero
        std::vector<int32_t> h_drow((size_t)B * ntiles * BT_TR, dummy), h_sid((size_t)B * NS, 0),
            h_sptr((size_t)B * (NS + 1), 0), h_cnt((size_t)B * 2, 0);
        for (int64_t p = 0; p < B; p++) {
            for (size_t k = 0; k < drows[(size_t)p].size(); k++) h_drow[(size_t)p * ntiles * BT_TR + k] = drows[(size_t)p][k];
            int32_t acc = 0;
            for (size_t k = 0; k < srows[(size_t)p].size(); k++) {
                h_sid[(size_t)p * NS + k] = srows[(size_t)p][k];
                h_sptr[(size_t)p * (NS + 1) + k] = acc;
                acc += nnz[(size_t)p * m + srows[(size_t)p][k]];
            }
            for (size_t k = srows[(size_t)p].size(); k <= (size_t)NS; k++) h_sptr[(size_t)p * (NS + 1) + k] = acc;
            h_cnt[(size_t)p * 2] = (int32_t)srows[(size_t)p].size();
        }
        d_drow.upload(h_drow);
        d_sid.upload(h_sid);
        d_sptr.upload(h_sptr);
        dA.alloc((size_t)B * a_stride, false);
        d_scol.alloc((size_t)B * NZ);
        d_sval.alloc((size_t)B * NZ);
        for (int64_t p0 = 0; p0 < B; p0 += 32768) {
            const int64_t cnt = std::min<int64_t>(32768, B - p0);
            k_gather_dense_rows<<<dim3((unsigned)(ntiles * BT_TR), (unsigned)cnt), 128, 0, h->stream>>>(
                src + (size_t)p0 * pstride_src, lda_src, pstride_src, m, n, d_drow.p + (size_t)p0 * ntiles * BT_TR,
                ntiles * BT_TR, dA.p + (size_t)p0 * a_stride, lda, a_stride);
            k_extract_sparse_rows<<<dim3((unsigned)NS, (unsigned)cnt), 32, 0, h->stream>>>(
                src + (size_t)p0 * pstride_src, lda_src, pstride_src, n, d_sid.p + (size_t)p0 * NS,
                d_sptr.p + (size_t)p0 * (NS + 1), NS, d_scol.p + (size_t)p0 * NZ, d_sval.p + (size_t)p0 * NZ, NZ);
        }
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        // CSC of the sparse rows, on the host (sparse_nnz_total entries)
        std::vector<int32_t> h_scol((size_t)B * NZ);
        std::vector<double> h_sval((size_t)B * NZ);
        FOS_CUDA(cudaMemcpy(h_scol.data(), d_scol.p, h_scol.size() * 4, cudaMemcpyDeviceToHost));
        FOS_CUDA(cudaMemcpy(h_sval.data(), d_sval.p, h_sval.size() * 8, cudaMemcpyDeviceToHost));
        std::vector<std::vector<int32_t>> ccols((size_t)B);
        int64_t NCSmax = 0;
        std::vector<int32_t> colcnt((size_t)n);
        for (int64_t p = 0; p < B; p++) {
            std::fill(colcnt.begin(), colcnt.end(), 0);
            const int32_t nzp = h_sptr[(size_t)p * (NS + 1) + NS];
            for (int32_t q = 0; q < nzp; q++) colcnt[(size_t)h_scol[(size_t)p * NZ + q]]++;
            for (int64_t j = 0; j < n; j++)
                if (colcnt[(size_t)j]) ccols[(size_t)p].push_back((int32_t)j);
            NCSmax = std::max<int64_t>(NCSmax, (int64_t)ccols[(size_t)p].size());
        }
        NCS = (int32_t)std::max<int64_t>(NCSmax, 1);
        std::vector<int32_t> h_cid((size_t)B * NCS, 0), h_cptr((size_t)B * (NCS + 1), 0), h_crow((size_t)B * NZ, 0);
        std::vector<double> h_cval((size_t)B * NZ, 0.0);
        std::vector<int32_t> slot((size_t)n), fill;
        for (int64_t p = 0; p < B; p++) {
            const std::vector<int32_t> &cc = ccols[(size_t)p];
            std::fill(colcnt.begin(), colcnt.end(), 0);
            const int32_t nzp = h_sptr[(size_t)p * (NS + 1) + NS];
            for (int32_t q = 0; q < nzp; q++) colcnt[(size_t)h_scol[(size_t)p * NZ + q]]++;
            int32_t acc = 0;
            fill.assign(cc.size(), 0);
            for (size_t k = 0; k < cc.size(); k++) {
                slot[(size_t)cc[k]] = (int32_t)k;
                h_cid[(size_t)p * NCS + k] = cc[k];
                h_cptr[(size_t)p * (NCS + 1) + k] = acc;
                fill[k] = acc;
                acc += colcnt[(size_t)cc[k]];
            }
            for (size_t k = cc.size(); k <= (size_t)NCS; k++) h_cptr[(size_t)p * (NCS + 1) + k] = acc;
            for (int32_t sr = 0; sr < h_cnt[(size_t)p * 2]; sr++)  // row order inside a column = sparse-row order
                for (int32_t q = h_sptr[(size_t)p * (NS + 1) + sr]; q < h_sptr[(size_t)p * (NS + 1) + sr + 1]; q++) {
                    const int32_t k = slot[(size_t)h_scol[(size_t)p * NZ + q]];
                    const int32_t pos = fill[(size_t)k]++;
                    h_crow[(size_t)p * NZ + pos] = h_sid[(size_t)p * NS + sr];
                    h_cval[(size_t)p * NZ + pos] = h_sval[(size_t)p * NZ + q];
                }
            h_cnt[(size_t)p * 2 + 1] = (int32_t)cc.size();
        }
        d_cnt.upload(h_cnt);
        d_cid.upload(h_cid);
        d_cptr.upload(h_cptr);
        d_crow.upload(h_crow);
        d_cval.upload(h_cval);
        h->stats.launches += 3;
    }
    // b, c and their norms (host vectors)
    {
        std::vector<double> hb((size_t)B * L.m_pad, 0.0), hc((size_t)B * L.n_pad, 0.0), hnb((size_t)B), hnc((size_t)B);
        for (int64_t p = 0; p < B; p++) {
            double sb = 0, sc = 0;
            for (int64_t i = 0; i < m; i++) {
                const double v = b[(size_t)p * m + i];
                hb[(size_t)p * L.m_pad + i] = v;
                sb += v * v;
            }
            for (int64_t j = 0; j < n; j++) {
                const double v = c[(size_t)p * n + j];
                hc[(size_t)p * L.n_pad + j] = v;
                sc += v * v;
            }
            hnb[(size_t)p] = std::sqrt(sb);
            hnc[(size_t)p] = std::sqrt(sc);
        }
        db.upload(hb);
        dc.upload(hc);
        dnb.upload(hnb);
        dncn.upload(hnc);
    }
    // cones: DualConeProduct (cones.jl:114-142), same construction as Handle::load_conic
    {
        std::vector<ConeSeg> segs;
        auto add = [&](int64_t base, int64_t nc, const int32_t *t, const int64_t *ln, int dual) {
            int64_t off = base;
            for (int64_t k = 0; k < nc; k++) {
                FOS_REQUIRE(ln[k] >= 0, "negative cone length");
                if (t[k] == FOS_CONE_SDP || t[k] == FOS_CONE_SOCROT || t[k] == FOS_CONE_EXPPRIMAL || t[k] == FOS_CONE_EXPDUAL)
                    throw Error(FOS_ERR_UNSUPPORTED, "batch mode offers Free, Zero, NonNeg, NonPos and SOC cones");
                segs.push_back(ConeSeg{t[k], dual, off, ln[k]});
                off += ln[k];
            }
        };
        add(0, nc2, t2, l2, 0);
        add(L.n_pad, nc1, t1, l1, 1);
        segs.push_back(ConeSeg{FOS_CONE_NONNEG, 0, L.n_pad + L.m_pad, 1});
        add(L.LP, nc2, t2, l2, 1);
        add(L.LP + L.n_pad, nc1, t1, l1, 0);
        segs.push_back(ConeSeg{FOS_CONE_NONNEG, 0, L.LP + L.n_pad + L.m_pad, 1});
        cones.build(L.NP, segs);
        if (cones.nsoc > BT_MAX_SOC)
            throw Error(FOS_ERR_UNSUPPORTED, "batch mode supports at most " + std::to_string(BT_MAX_SOC / 2) + " SOC cones");
    }
    // CG tolerance schedule max(0.2^sqrt(i), l*eps) (affinepluslinear.jl:108-112), tabulated on the host so
    // that the device uses bit-identical values to the single-problem path
    {
        tol_floor = (double)l * 2.220446049250313e-16;
        std::vector<double> tt;
        for (int64_t i = 1; i < 100000; i++) {
            const double t = std::pow(0.2, std::sqrt((double)i));
            if (t <= tol_floor) break;
            tt.push_back(t);
        }
        tol_n = (int)tt.size();
        if (tt.empty()) tt.push_back(tol_floor);
        dtol.upload(tt);
    }
    vec.alloc((size_t)B * BV_COUNT * L.NP);
    dctl.alloc((size_t)B);
    counter.alloc(1);
    ensure_recs(128);
    k_batch_reset<<<(unsigned)((B + 127) / 128), 128, 0, h->stream>>>(dctl.p, (int)B, 0);
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    loaded = true;
    set_initial_iterate();
}

void BatchSolver::ensure_recs(int cap)
{
    if (cap <= rec_cap) return;
    drecs.alloc((size_t)B * cap * FOS_REC_LEN);
    rec_cap = cap;
}

void BatchSolver::set_algorithm()
{
    // init_algorithm! (gap.jl:23-28 etc.): alpha12 = 2, t = 1, p = q = 0; S1 keeps its state
    FOS_REQUIRE(h->alg != FOS_ALG_GAPP, "GAPP is not offered in batch mode");
    k_batch_reset<<<(unsigned)((B + 127) / 128), 128, 0, h->stream>>>(dctl.p, (int)B, 2);
    for (int which : {BV_FY, BV_FXOLD, BV_DP, BV_DQ})
        FOS_CUDA(cudaMemset2DAsync(vec.p + (size_t)which * L.NP, (size_t)BV_COUNT * L.NP * 8, 0, (size_t)L.NP * 8,
                                   (size_t)B, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
}

void BatchSolver::begin_solve()
{
    k_batch_reset<<<(unsigned)((B + 127) / 128), 128, 0, h->stream>>>(dctl.p, (int)B, 1);
}

void BatchSolver::set_initial_iterate()
{
    // HSDE_getinitialvalue (HSDE.jl:40-47): zeros except tau = kappa = 1
    FOS_CUDA(cudaMemset2DAsync(vec.p + (size_t)BV_X * L.NP, (size_t)BV_COUNT * L.NP * 8, 0, (size_t)L.NP * 8, (size_t)B,
                               h->stream));
    const double one = 1.0;
    std::vector<double> ones((size_t)B, one);
    const int64_t ot = L.n_pad + L.m_pad;
    FOS_CUDA(cudaMemcpy2DAsync(vec.p + (size_t)BV_X * L.NP + ot, (size_t)BV_COUNT * L.NP * 8, ones.data(), 8, 8, (size_t)B,
                               cudaMemcpyHostToDevice, h->stream));
    FOS_CUDA(cudaMemcpy2DAsync(vec.p + (size_t)BV_X * L.NP + L.LP + ot, (size_t)BV_COUNT * L.NP * 8, ones.data(), 8, 8,
                               (size_t)B, cudaMemcpyHostToDevice, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
}

void BatchSolver::set_vector(int which, const double *z, int64_t pb0, int64_t count)
{
    FOS_REQUIRE(which >= 0 && which < BV_COUNT && pb0 >= 0 && count >= 1 && pb0 + count <= B, "bad batch vector range");
    DevBuf<double> st;
    st.alloc((size_t)count * N, false);
    FOS_CUDA(cudaMemcpyAsync(st.p, z, (size_t)count * N * 8, cudaMemcpyHostToDevice, h->stream));
    dim3 g((unsigned)std::min<int64_t>((L.NP + VBLOCK - 1) / VBLOCK, 64), (unsigned)count);
    k_pack_batch<<<g, VBLOCK, 0, h->stream>>>(seg, st.p, N, vec.p + ((size_t)pb0 * BV_COUNT + which) * L.NP, L.NP,
                                               (int64_t)BV_COUNT * L.NP);
    FOS_CUDA(cudaStreamSynchronize(h->stream));
}

void BatchSolver::get_vector(int which, double *z, int64_t pb0, int64_t count)
{
    FOS_REQUIRE(which >= 0 && which < BV_COUNT && pb0 >= 0 && count >= 1 && pb0 + count <= B, "bad batch vector range");
    DevBuf<double> st;
    st.alloc((size_t)count * N, false);
    dim3 g((unsigned)std::min<int64_t>((N + VBLOCK - 1) / VBLOCK, 64), (unsigned)count);
    k_unpack_batch<<<g, VBLOCK, 0, h->stream>>>(seg, vec.p + ((size_t)pb0 * BV_COUNT + which) * L.NP,
                                                 (int64_t)BV_COUNT * L.NP, st.p, N);
    FOS_CUDA(cudaMemcpyAsync(z, st.p, (size_t)count * N * 8, cudaMemcpyDeviceToHost, h->stream));
    FOS_CUDA(cudaStreamSynchronize(h->stream));
}

void BatchSolver::launch(int64_t i_start, int64_t n_iters, int64_t checki, double eps, bool do_run, bool do_finish)
{
    FOS_REQUIRE(loaded, "no batch loaded on this handle");
    FOS_REQUIRE(checki >= 1 && n_iters >= 0, "bad run arguments");
    FOS_REQUIRE(h->alg != FOS_ALG_GAPP, "GAPP is not offered in batch mode");
    ensure_recs((int)std::min<int64_t>(n_iters / checki + 2, 1 << 16));
    if (do_run) last_eps = eps;
    BatchArgs a;
    a.L = L;
    a.lda = lda;
    a.ntiles = ntiles;
    a.S = S;
    a.CW = CW;
    a.KP = KP;
    a.A = dA.p;
    a.a_stride = a_stride;
    a.drow = d_drow.p;
    a.sp_count = d_cnt.p;
    a.srow_ptr = d_sptr.p;
    a.srow_id = d_sid.p;
    a.scol = d_scol.p;
    a.sval = d_sval.p;
    a.ccol_ptr = d_cptr.p;
    a.ccol_id = d_cid.p;
    a.crow = d_crow.p;
    a.cval = d_cval.p;
    a.NS = NS;
    a.NCS = NCS;
    a.NZ = NZ;
    a.mr = (int32_t)(L.m_pad + 16);
    a.b = db.p;
    a.c = dc.p;
    a.nb = dnb.p;
    a.ncn = dncn.p;
    a.vec = vec.p;
    a.ctl = dctl.p;
    a.recs = drecs.p;
    a.rec_cap = rec_cap;
    a.tol_table = dtol.p;
    a.tol_n = tol_n;
    a.tol_floor = tol_floor;
    a.ops = cones.ops.p;
    a.cone_of = cones.cone_of.p;
    a.soc = cones.soc.p;
    a.nsoc = cones.nsoc;
    a.work_counter = counter.p;
    a.B = (int)B;
    a.alg = h->alg;
    a.alpha = h->alpha;
    a.alpha1 = h->alpha1;
    a.alpha2 = h->alpha2;
    a.betaA = h->betaA;
    a.i_start = i_start;
    a.n_iters = n_iters;
    a.checki = checki;
    a.eps = eps;
    a.do_run = do_run ? 1 : 0;
    a.do_finish = do_finish ? 1 : 0;
    FOS_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned int), h->stream));
    const int grid = (int)std::min<int64_t>(B, grid_ctas > 0 ? grid_ctas : ctas_per_sm * h->num_sms);
    const int threads = (CW + 1) * 32;
#define BT_LAUNCH(KPV, MAXT, MINB)                                                                          \
    FOS_CUDA(cudaFuncSetAttribute(k_batch_solve<KPV, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)smem_bytes));                                                         \
    k_batch_solve<KPV, MAXT, MINB><<<grid, threads, smem_bytes, h->stream>>>(a);
    if (ctas_per_sm == 2) {
        BT_LAUNCH(1, 320, 2)
    } else {
        switch (KP) {
        case 1: BT_LAUNCH(1, 512, 1) break;
        case 2: BT_LAUNCH(2, 512, 1) break;
        case 3: BT_LAUNCH(3, 512, 1) break;
        default: BT_LAUNCH(4, 512, 1) break;
        }
    }
#undef BT_LAUNCH
    h->stats.launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(FOS_ERR_CUDA, std::string("batch kernel launch failed: ") + cudaGetErrorString(e));
}

void BatchSolver::collect(int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap_host, int64_t *n_rec,
                          int64_t *cgiter_total)
{
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<BatchCtl> hc((size_t)B);
    FOS_CUDA(cudaMemcpy(hc.data(), dctl.p, (size_t)B * sizeof(BatchCtl), cudaMemcpyDeviceToHost));
    std::vector<double> hr;
    if (records && rec_cap_host > 0) {
        hr.resize((size_t)B * rec_cap * FOS_REC_LEN);
        FOS_CUDA(cudaMemcpy(hr.data(), drecs.p, hr.size() * 8, cudaMemcpyDeviceToHost));
    }
    for (int64_t p = 0; p < B; p++) {
        const BatchCtl &c = hc[(size_t)p];
        if (iters_done) iters_done[p] = c.iters_done;
        if (status) status[p] = c.status;
        if (n_rec) n_rec[p] = c.nrec;
        if (cgiter_total) cgiter_total[p] = c.total_cg;
        if (c.warn_maxit) h->warn_maxit = true;
        if (records && rec_cap_host > 0) {
            const int64_t ncopy = std::min<int64_t>(std::min<int64_t>(c.nrec, rec_cap_host), rec_cap);
            if (ncopy > 0)
                memcpy(records + (size_t)p * rec_cap_host * FOS_REC_LEN, hr.data() + (size_t)p * rec_cap * FOS_REC_LEN,
                       (size_t)ncopy * FOS_REC_LEN * 8);
        }
    }
}

// algorithmic bytes of one pass over one problem's A, batch average: dense rows as FP64 + sparse rows as
// CSR and CSC entries (8-byte value + 4-byte index each)
double BatchSolver::bytes_per_pass() const
{
    return (8.0 * (double)dense_rows_total * (double)L.n + 24.0 * (double)sparse_nnz_total) / (double)B;
}

}  // namespace fos
