// kernels.cuh -- K2 (Q / KKT epilogues fused with the CG dot products), K3 (CG vector
// updates with fused reductions), K4 (cone projections fused with the relaxation /
// averaging steps of GAP, GAPA, FISTA, Dykstra, GAPP) and K6 (residual check).
//
// All vectors use the padded device layout (see solver.cuh): every segment starts on a
// 128-byte boundary and its padding holds zeros, so each kernel is a flat loop.
// Arithmetic that the reference performs as separate broadcast operations is done with
// non-contracting intrinsics (mul_/add_/sub_) in the same order.
#pragma once
#include "common.cuh"

namespace fos {

constexpr int VBLOCK = 256;  // block size of every vector kernel

struct Lay {
    int32_t form;  // 0 = HSDE conic (operator Q, l x l), 1 = plain matrix operator (am x an), 2 = symmetric n x n
    int64_t n, m;  // A is m x n
    int64_t n_pad, m_pad;
    int64_t LP;  // HSDE: padded half length n_pad + m_pad + PAD.  plain: unused (0)
    int64_t NP;  // padded iterate length
    // HSDE offsets inside a half: x at 0, y at n_pad, tau at n_pad + m_pad
};

enum { K2_OUT = 0, K2_AP = 1, K2_RESID = 2 };

#ifdef __CUDACC__

__device__ __forceinline__ bool cg_skip(const Ctrl *ctrl) { return ctrl->done != 0; }

// =======================================================================================
// K2, HSDE operator:  out = [I Q'; Q -I] * in     (affinepluslinear.jl:37-49 on top of
// HSDEAffine.jl:41-65).  V holds A*[in1.x in2.x] (v = 0,1) and A'*[in1.y in2.y].
//   K2_OUT   : out written.
//   K2_AP    : in = p, out = Ap, fused <Ap,p>; last block: alpha = rn / <Ap,p>   (cg :38-39)
//   K2_RESID : in = x0, r = rhs - out, p = r, fused <r,r>; last block: rn, iter = 1 (cg :32-36)
// =======================================================================================
template <int MODE>
__global__ void __launch_bounds__(VBLOCK)
k2_kkt_hsde(Lay L, MVView V, const double *__restrict__ in, const double *__restrict__ c,
            const double *__restrict__ b, double *__restrict__ out, const double *__restrict__ rhs,
            double *__restrict__ r, double *__restrict__ p, Ctrl *ctrl, RedBuf rb, int predicated)
{
    if (predicated && cg_skip(ctrl)) return;
    const int64_t LP = L.LP, oy = L.n_pad, ot = L.n_pad + L.m_pad;
    const double tau1 = in[ot], tau2 = in[LP + ot];
    double q[5] = {0, 0, 0, 0, 0};  // c.X1, b.W1, c.X2, b.W2, fused dot
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < ot; e += (int64_t)gridDim.x * VBLOCK) {
        double o1 = 0.0, o2 = 0.0;
        const double i1 = in[e], i2 = in[LP + e];
        if (e < oy) {
            if (e < L.n) {
                const double cj = c[e];
                // (Q B).x = A'B.y + B.tau*c   (HSDEAffine.jl:51,54)
                double w0, w1;
                mv_atw2(V, e, w0, w1);
                const double q1 = add_(w0, mul_(tau1, cj));
                const double q2 = add_(w1, mul_(tau2, cj));
                o1 = add_(-q2, i1);  // Q'in2 + in1  (transpose = negate, HSDEAffine.jl:61-65)
                o2 = sub_(q1, i2);   // Q in1 - in2
                q[0] = fma(cj, i1, q[0]);
                q[2] = fma(cj, i2, q[2]);
            }
        } else {
            const int64_t i = e - oy;
            if (i < L.m) {
                const double bi = b[i];
                // (Q B).y = -(A B.x - B.tau*b)   (HSDEAffine.jl:52,55,56)
                double a0, a1;
                mv_ax2(V, i, a0, a1);
                const double q1 = -sub_(a0, mul_(tau1, bi));
                const double q2 = -sub_(a1, mul_(tau2, bi));
                o1 = add_(-q2, i1);
                o2 = sub_(q1, i2);
                q[1] = fma(bi, i1, q[1]);
                q[3] = fma(bi, i2, q[3]);
            }
        }
        if (MODE == K2_OUT) {
            out[e] = o1;
            out[LP + e] = o2;
        } else if (MODE == K2_AP) {
            out[e] = o1;
            out[LP + e] = o2;
            q[4] = fma(o1, i1, q[4]);
            q[4] = fma(o2, i2, q[4]);
        } else {
            const double r1 = sub_(rhs[e], o1), r2 = sub_(rhs[LP + e], o2);
            r[e] = r1;
            r[LP + e] = r2;
            p[e] = r1;
            p[LP + e] = r2;
            q[4] = fma(r1, r1, q[4]);
            q[4] = fma(r2, r2, q[4]);
        }
    }
    double tot[5];
    if (grid_reduce<5, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        // (Q B).tau = -c.B.x - b.B.y   (HSDEAffine.jl:57)
        const double q1t = sub_(-tot[0], tot[1]);
        const double q2t = sub_(-tot[2], tot[3]);
        const double o1 = add_(-q2t, tau1);
        const double o2 = sub_(q1t, tau2);
        if (MODE == K2_OUT) {
            out[ot] = o1;
            out[LP + ot] = o2;
        } else if (MODE == K2_AP) {
            out[ot] = o1;
            out[LP + ot] = o2;
            const double pAp = tot[4] + o1 * tau1 + o2 * tau2;
            ctrl->alpha = ctrl->rn / pAp;
        } else {
            const double r1 = sub_(rhs[ot], o1), r2 = sub_(rhs[LP + ot], o2);
            r[ot] = r1;
            r[LP + ot] = r2;
            p[ot] = r1;
            p[LP + ot] = r2;
            ctrl->rn = tot[4] + r1 * r1 + r2 * r2;
            ctrl->iter = 1;
        }
    }
}

// K2, plain operator:  out = [I A'; A -I] * in   with V = (A*in1, A'*in2), one right-hand side.
template <int MODE>
__global__ void __launch_bounds__(VBLOCK)
k2_kkt_plain(Lay L, MVView V, const double *__restrict__ in, double *__restrict__ out,
             const double *__restrict__ rhs, double *__restrict__ r, double *__restrict__ p, Ctrl *ctrl, RedBuf rb,
             int predicated)
{
    if (predicated && cg_skip(ctrl)) return;
    double q[1] = {0};
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < L.NP; e += (int64_t)gridDim.x * VBLOCK) {
        double o = 0.0;
        const double ie = in[e];
        if (e < L.n_pad) {
            if (e < L.n) o = add_(mv_atw(V, 0, e), ie);  // affinepluslinear.jl:45-46
        } else {
            const int64_t i = e - L.n_pad;
            if (i < L.m) o = sub_(mv_ax(V, 0, i), ie);  // :47-48
        }
        if (MODE == K2_OUT) {
            out[e] = o;
        } else if (MODE == K2_AP) {
            out[e] = o;
            q[0] = fma(o, ie, q[0]);
        } else {
            const double re = sub_(rhs[e], o);
            r[e] = re;
            p[e] = re;
            q[0] = fma(re, re, q[0]);
        }
    }
    double tot[1];
    if (grid_reduce<1, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        if (MODE == K2_AP) ctrl->alpha = ctrl->rn / tot[0];
        else if (MODE == K2_RESID) {
            ctrl->rn = tot[0];
            ctrl->iter = 1;
        }
    }
}

// K2, symmetric matrix operator (form 2): out = A*in.  Serves conjugategradient! on a plain SPD
// matrix (test/conjugateGradient.jl) through fos_cg_dense.
template <int MODE>
__global__ void __launch_bounds__(VBLOCK)
k2_spd(Lay L, MVView V, const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ rhs,
       double *__restrict__ r, double *__restrict__ p, Ctrl *ctrl, RedBuf rb, int predicated)
{
    if (predicated && cg_skip(ctrl)) return;
    double q[1] = {0};
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < L.NP; e += (int64_t)gridDim.x * VBLOCK) {
        const double o = e < L.n ? mv_ax(V, 0, e) : 0.0;
        if (MODE == K2_OUT) {
            out[e] = o;
        } else if (MODE == K2_AP) {
            out[e] = o;
            q[0] = fma(o, in[e], q[0]);
        } else {
            const double re = sub_(rhs[e], o);
            r[e] = re;
            p[e] = re;
            q[0] = fma(re, re, q[0]);
        }
    }
    double tot[1];
    if (grid_reduce<1, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        if (MODE == K2_AP) ctrl->alpha = ctrl->rn / tot[0];
        else if (MODE == K2_RESID) {
            ctrl->rn = tot[0];
            ctrl->iter = 1;
        }
    }
}

// Y = Q*B or Q'*B (HSDEAffine.jl:41-65) from V = (A*B.x, A'*B.y); unit-level entry point.
static __global__ void __launch_bounds__(VBLOCK)
k2_q_hsde(Lay L, MVView V, const double *__restrict__ B, const double *__restrict__ c, const double *__restrict__ b,
          double *__restrict__ Y, int transpose, RedBuf rb)
{
    const int64_t oy = L.n_pad, ot = L.n_pad + L.m_pad;
    const double tau = B[ot];
    double q[2] = {0, 0};
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < ot; e += (int64_t)gridDim.x * VBLOCK) {
        double o = 0.0;
        if (e < oy) {
            if (e < L.n) {
                o = add_(mv_atw(V, 0, e), mul_(tau, c[e]));
                q[0] = fma(c[e], B[e], q[0]);
            }
        } else {
            const int64_t i = e - oy;
            if (i < L.m) {
                o = -sub_(mv_ax(V, 0, i), mul_(tau, b[i]));
                q[1] = fma(b[i], B[e], q[1]);
            }
        }
        Y[e] = transpose ? -o : o;
    }
    double tot[2];
    if (grid_reduce<2, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        const double o = sub_(-tot[0], tot[1]);
        Y[ot] = transpose ? -o : o;
    }
}

// rhs of the affine projection (affinepluslinear.jl:94-95):  rhs1 = beta*Op'x2 + x1 - q.
// HSDE (beta = 1, q = 0, b = 0; HSDE.jl:22): V = (A*x2.x, A'*x2.y), x2 = (r, s, kappa).
static __global__ void __launch_bounds__(VBLOCK)
k2_rhs_hsde(Lay L, MVView V, const double *__restrict__ xin, const double *__restrict__ c,
            const double *__restrict__ b, double *__restrict__ rhs, RedBuf rb)
{
    const int64_t LP = L.LP, oy = L.n_pad, ot = L.n_pad + L.m_pad;
    const double kap = xin[LP + ot];
    double q[2] = {0, 0};
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < ot; e += (int64_t)gridDim.x * VBLOCK) {
        double o = 0.0;
        const double x2 = xin[LP + e];
        if (e < oy) {
            if (e < L.n) {
                const double t = -add_(mv_atw(V, 0, e), mul_(kap, c[e]));  // (Q'x2).x
                o = add_(t, xin[e]);
                q[0] = fma(c[e], x2, q[0]);
            }
        } else {
            const int64_t i = e - oy;
            if (i < L.m) {
                const double t = sub_(mv_ax(V, 0, i), mul_(kap, b[i]));  // (Q'x2).y = -(Q x2).y
                o = add_(t, xin[e]);
                q[1] = fma(b[i], x2, q[1]);
            }
        }
        rhs[e] = o;
    }
    double tot[2];
    if (grid_reduce<2, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        const double t = -sub_(-tot[0], tot[1]);
        rhs[ot] = add_(t, xin[ot]);
    }
}

// plain operator: rhs1 = beta*A'x2 + x1 - q ; rhs2 keeps b (pre-stored, affinepluslinear.jl:76-77)
static __global__ void __launch_bounds__(VBLOCK)
k2_rhs_plain(Lay L, MVView V, const double *__restrict__ xin, const double *__restrict__ qv, double beta,
             double *__restrict__ rhs)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < L.n; e += (int64_t)gridDim.x * VBLOCK)
        rhs[e] = sub_(add_(mul_(beta, mv_atw(V, 0, e)), xin[e]), qv[e]);
}

// Fused right-hand side (option "fuse_rhs"): by linearity the initial CG residual
//     r0 = rhs - KKT*x0,   rhs = [beta*Op'x2 + x1 - q ; b]      (affinepluslinear.jl:94-95, cg :32-33)
// equals   rhs' - KKT*d   with  d = [x0_1 ; x0_2 - beta*x2]  and  rhs' = [x1 - q ; b + beta*x2],
// which needs ONE pass over A instead of two (the rhs pass disappears).  HSDE: q = b = 0, beta = 1,
// so rhs' is the input vector itself.  d is written to `dvec`; for the plain operator rhs' to `rhsp`.
static __global__ void __launch_bounds__(VBLOCK)
k_fuse_prep(Lay L, const double *__restrict__ x0, const double *__restrict__ xin, double beta,
            const double *__restrict__ qv, const double *__restrict__ bhat, double *__restrict__ dvec,
            double *__restrict__ rhsp)
{
    const int64_t p2 = L.form == 0 ? L.LP : L.n_pad;
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < L.NP; e += (int64_t)gridDim.x * VBLOCK) {
        if (e < p2) {
            dvec[e] = x0[e];
            if (L.form == 1) rhsp[e] = sub_(xin[e], qv[e]);
        } else {
            const double bx2 = mul_(beta, xin[e]);
            dvec[e] = sub_(x0[e], bx2);
            if (L.form == 1) rhsp[e] = add_(bhat[e - p2], bx2);
        }
    }
}

// =======================================================================================
// K3: CG updates (conjugategradients.jl:40-51)
// =======================================================================================
// x += alpha p ; r -= alpha Ap ; ||r|| ; stop test ; beta
static __global__ void __launch_bounds__(VBLOCK)
k3_cg_update(int64_t NP, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
             const double *__restrict__ Ap, Ctrl *ctrl, RedBuf rb)
{
    if (cg_skip(ctrl)) return;
    const double alpha = ctrl->alpha;
    double q[1] = {0};
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        x[e] = add_(x[e], mul_(alpha, p[e]));               // :40
        const double re = sub_(r[e], mul_(alpha, Ap[e]));   // :41
        r[e] = re;
        q[0] = fma(re, re, q[0]);
    }
    double tot[1];
    if (grid_reduce<1, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        const double rr = tot[0];
        const double rnorm = sqrt(rr);
        ctrl->rnorm = rnorm;
        if (rnorm <= ctrl->tol || ctrl->iter >= ctrl->max_iters) {  // :42
            ctrl->done = 1;
            if (ctrl->iter >= ctrl->max_iters) ctrl->warn_maxit = 1;  // :53
        } else {
            const double rnold = ctrl->rn;  // :45
            ctrl->rn = rr;                  // :46
            ctrl->beta = rr / rnold;        // :47
            ctrl->iter += 1;                // :51
        }
    }
}
// p = beta p + r   (:49-50, two roundings)
static __global__ void __launch_bounds__(VBLOCK)
k3_cg_dir(int64_t NP, double *__restrict__ p, const double *__restrict__ r, const Ctrl *ctrl)
{
    if (cg_skip(ctrl)) return;
    const double beta = ctrl->beta;
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK)
        p[e] = add_(mul_(beta, p[e]), r[e]);
}

static __global__ void k_cg_begin(Ctrl *ctrl, double tol, int max_iters)
{
    ctrl->tol = tol;
    ctrl->max_iters = max_iters;
    ctrl->done = 0;
    ctrl->iter = 0;
}

// Graph path, first kernel of an outer iteration on the conic form (Handle::run_graph): k_fuse_prep + k_cg_begin in
// one launch, with the per-iteration scalars taken from the DEVICE so that the captured graph never changes:
//   dvec = [x0_1 ; x0_2 - x_2]                              (k_fuse_prep with beta = 1, q = b = 0; HSDE.jl:22)
//   tol  = tol_table[S1.i] ; S1.i += 1                      (affinepluslinear.jl:108-114; the table holds the host's
//                                                            max(0.2^sqrt(i), l*eps) so that both paths compare ||r||
//                                                            with the same bits)
//   i += 1 ; done = 0 ; iter = 0
//   FISTA: told = t ; t = (1 + sqrt(1 + 4 told^2)) / 2 ; coef = (told - 1) / t   (fista.jl:44-46; every operation
//   rounded separately, like the host's -- the legacy path computes the same bits in Handle::step)
static __global__ void __launch_bounds__(VBLOCK)
k_iter_begin(Lay L, const double *__restrict__ x0, const double *__restrict__ xin, double *__restrict__ dvec,
             Ctrl *ctrl, const double *__restrict__ tol_table, int tbl_n, int max_iters, int fista)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < L.NP; e += (int64_t)gridDim.x * VBLOCK)
        dvec[e] = e < L.LP ? x0[e] : sub_(x0[e], xin[e]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t k = ctrl->s1_calls_dev;
        ctrl->tol = tol_table[k < tbl_n ? k : tbl_n - 1];
        ctrl->s1_calls_dev = k + 1;
        ctrl->cur_i += 1;
        ctrl->max_iters = max_iters;
        ctrl->done = 0;
        ctrl->iter = 0;
        if (fista) {
            const double told = ctrl->fista_t;
            const double t = __ddiv_rn(__dadd_rn(1.0, __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(__dmul_rn(4.0, told), told)))), 2.0);
            ctrl->fista_t = t;
            ctrl->fista_coef = __ddiv_rn(__dsub_rn(told, 1.0), t);
        }
    }
}

// =======================================================================================
// generic vector helpers
// =======================================================================================
// out = a*X + b*Y with three roundings (the reference's  y .= a.*y .+ (1-a).*x  form).
// part2_scale multiplies X on entries >= part2_off first (the y2 .*= beta of affinepluslinear.jl:124).
// use_a12: take a from ctrl->alpha12 and b = 1 - a (GAPA).
static __global__ void __launch_bounds__(VBLOCK)
k_relax(int64_t NP, double *__restrict__ out, double a, const double *__restrict__ X, double bcoef,
        const double *__restrict__ Y, int64_t part2_off, double part2_scale, const Ctrl *ctrl, int use_a12)
{
    if (use_a12) {
        a = ctrl->alpha12;
        bcoef = 1.0 - a;
    }
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        double xe = X[e];
        if (part2_scale != 1.0 && e >= part2_off) xe = mul_(xe, part2_scale);
        out[e] = add_(mul_(a, xe), mul_(bcoef, Y[e]));
    }
}
// out = X (with the part-2 scaling)
static __global__ void __launch_bounds__(VBLOCK)
k_copy_scaled(int64_t NP, double *__restrict__ out, const double *__restrict__ X, int64_t part2_off,
              double part2_scale)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        double xe = X[e];
        if (part2_scale != 1.0 && e >= part2_off) xe = mul_(xe, part2_scale);
        out[e] = xe;
    }
}
// out = X + a*Y
static __global__ void __launch_bounds__(VBLOCK)
k_add_scaled(int64_t NP, double *out, const double *X, double a, const double *Y, const Ctrl *ctrl, int a_from_ls)
{
    if (a_from_ls) a = ctrl->ls_alphabest;
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK)
        out[e] = add_(X[e], mul_(a, Y[e]));
}
// out = X - Y
static __global__ void __launch_bounds__(VBLOCK)
k_sub(int64_t NP, double *out, const double *X, const double *Y)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK)
        out[e] = sub_(X[e], Y[e]);
}
// Dykstra (dykstra.jl:29-31):  w = x + p  (input of P1)
// and after P1:  p = (x + p) - y   ;   w2 = y + q  (input of P2)
static __global__ void __launch_bounds__(VBLOCK)
k_dykstra_mid(int64_t NP, const double *__restrict__ w, const double *__restrict__ y, double *__restrict__ p,
              const double *__restrict__ q, double *__restrict__ w2)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        p[e] = sub_(w[e], y[e]);
        w2[e] = add_(y[e], q[e]);
    }
}

// =======================================================================================
// K4: cone projections (cones.jl) fused with the algorithm's relaxation / averaging
// =======================================================================================
enum : uint8_t {
    OP_ZERO = 0,      // Zero cone, dual of Free, padding
    OP_COPY = 1,      // Free cone, dual of Zero
    OP_MAX0 = 2,      // NonNeg (self-dual), tau, kappa
    OP_MIN0 = 3,      // NonPos (self-dual per cones.jl:102)
    OP_SOC_HEAD = 4,
    OP_SOC_TAIL = 5,
    OP_SOCD_HEAD = 6,  // dual via Moreau, cones.jl:80-85
    OP_SOCD_TAIL = 7,
    OP_PRE = 8,        // projected value already in proj[] (PSD cones, K5)
    OP_RSOC_H0 = 9,    // rotated SOC (IndRotatedSOC): first, second entry, tail
    OP_RSOC_H1 = 10,
    OP_RSOC_TAIL = 11,
    OP_RSOCD_H0 = 12,  // ... its dual via Moreau
    OP_RSOCD_H1 = 13,
    OP_RSOCD_TAIL = 14,
    OP_EXPP = 15,      // exponential cone (IndExpPrimal), 3 entries, cone_of[e] = position 0..2
    OP_EXPP_D = 16,    // proxDual! of ExpPrimal: x + P_K(-x)
    OP_EXPD = 17,      // IndExpDual: x + P_K(-x) as well (Moreau)
    OP_EXPD_D = 18,    // proxDual! of ExpDual: x + P_K*(-x)
    OP_BOX = 19        // IndBox(lo, hi): clamp, bounds in box[cone_of[e]]
};

struct SocCone {
    int64_t head;   // padded index of t
    int64_t len;    // total length including t
    int32_t chunk0; // first norm chunk
    int32_t nchunk;
    int32_t dual;
    int32_t rot;    // 1: rotated SOC, entries (x1, x2, w): the cone is 2 x1 x2 >= ||w||^2, x1, x2 >= 0
};
struct SocScale {  // result of the norm pass for one cone
    int32_t mode;  // primal: 0 -> 0, 1 -> copy, 2 -> scale.  dual: 0 -> copy x, 1 -> 0, 2 -> scale
    int32_t pad_;
    double rho;    // 0.5*(1 + t/nx) of the (possibly negated) input
    double nx;
};
constexpr int SOC_CHUNK = 2048;

constexpr double RSOC_C = 0.7071067811865475;  // sin(pi/4) = cos(pi/4), the literal of ProximalOperators' IndRotatedSOC
// rotated coordinates of the first two entries: x1' = c x1 + c x2, x2' = c x1 - c x2
__device__ __forceinline__ void rsoc_rotate(double a, double b, double &r1, double &r2)
{
    r1 = add_(mul_(RSOC_C, a), mul_(RSOC_C, b));
    r2 = sub_(mul_(RSOC_C, a), mul_(RSOC_C, b));
}
// From the squared tail norm: IndSOC on (t, w) / IndRotatedSOC on (x1, x2, w) (rotate the first two entries by
// pi/4, project onto the SOC with t = x1', tail (x2', w), rotate back).  Dual cones project -x (Moreau).
__device__ __forceinline__ SocScale soc_classify(const SocCone &K, double tail_sq, const double *__restrict__ in)
{
    double nx, t;
    const double sg = K.dual ? -1.0 : 1.0;
    if (K.rot) {
        double r1, r2;
        rsoc_rotate(sg * in[K.head], sg * in[K.head + 1], r1, r2);
        const double nw = sqrt(tail_sq);  // norm(x[3:end])
        nx = sqrt(add_(mul_(r2, r2), mul_(nw, nw)));
        t = r1;
    } else {
        nx = sqrt(tail_sq);
        t = sg * in[K.head];
    }
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
    return R;
}

// ---- exponential cone: cl{(r, s, t) : s > 0, s exp(r/s) <= t}.  ProximalOperators' IndExpPrimal follows the
// projection of SCS (cones.c, proj_exp_cone): membership tests, the analytic case r, s < 0, otherwise bisection on
// the dual variable rho with a 1-D Newton solve inside (tolerance 1e-8, at most 100 iterations each). ----
constexpr double EXP_TOL = 1e-8;
__device__ __forceinline__ double exp_newton_one_d(double rho, double y_hat, double z_hat)
{
    double t = fmax(-z_hat, 1e-6);
    for (int i = 0; i < 100; i++) {
        const double f = t * (t + z_hat) / rho / rho - y_hat / rho + log(t / rho) + 1.0;
        const double fp = (2.0 * t + z_hat) / rho / rho + 1.0 / t;
        t = t - f / fp;
        if (t <= -z_hat) return 0.0;
        else if (t <= 0.0) return z_hat;
        else if (fabs(f) < EXP_TOL) break;
    }
    return t + z_hat;
}
__device__ __forceinline__ double exp_calc_grad(const double (&v)[3], double (&x)[3], double rho)
{
    x[2] = exp_newton_one_d(rho, v[1], v[2]);
    x[1] = (x[2] - v[2]) * x[2] / rho;
    x[0] = v[0] - rho;
    if (x[1] <= 1e-12) return x[0];
    return x[0] + x[1] * log(x[1] / x[2]);
}
__device__ __forceinline__ void proj_exp_cone(const double (&v)[3], double (&y)[3])
{
    const double r = v[0], s = v[1], t = v[2];
    if ((s * exp(r / s) - t <= EXP_TOL && s > 0.0) || (r <= 0.0 && s == 0.0 && t >= 0.0)) {  // v in cl(K)
        y[0] = r; y[1] = s; y[2] = t;
        return;
    }
    if ((-r < 0.0 && r * exp(s / r) + exp(1.0) * t <= EXP_TOL) || (-r == 0.0 && -s >= 0.0 && -t >= 0.0)) {  // -v in K*
        y[0] = y[1] = y[2] = 0.0;
        return;
    }
    if (r < 0.0 && s < 0.0) {  // analytic
        y[0] = r; y[1] = 0.0; y[2] = fmax(t, 0.0);
        return;
    }
    double x[3], lb = 0.0, ub = 0.125;
    while (exp_calc_grad(v, x, ub) > 0.0) {
        lb = ub;
        ub *= 2.0;
    }
    for (int i = 0; i < 100; i++) {
        const double rho = (ub + lb) / 2.0;
        const double g = exp_calc_grad(v, x, rho);
        if (g > 0.0) lb = rho;
        else ub = rho;
        if (ub - lb < EXP_TOL) break;
    }
    y[0] = x[0]; y[1] = x[1]; y[2] = x[2];
}

// Relaxation fused in front of the cone kernels (graph path): the cone input is not read from memory but formed on the
// fly as  t1 = a*X + b*Y  (gap.jl:48 / gapa.jl:67 with a = alpha1 or alpha12, X = P1(x), Y = x; three roundings, like
// k_relax).  Both cone kernels evaluate the same expression, so they see identical bits; k4_cone_apply stores t1.
struct RelaxArgs {
    const double *X, *Y;
    double a, b;
    int use_a12;     // GAPA: a = ctrl->alpha12, b = 1 - a
    double *tmp1;    // where k4_cone_apply stores t1
};
template <bool RELAX>
__device__ __forceinline__ double relaxed_in(const double *__restrict__ in, const RelaxArgs &R, double a, double b,
                                             int64_t e)
{
    if (RELAX) return add_(mul_(a, R.X[e]), mul_(b, R.Y[e]));
    return in[e];
}

// pass 1: one block per chunk of a SOC tail, squared-norm partial; the last block folds the
// chunk partials per cone (in chunk order) and classifies each cone.
template <bool RELAX>
__global__ void __launch_bounds__(VBLOCK)
k4_soc_norms(const double *__restrict__ in, const SocCone *__restrict__ cones, int ncones,
             const int32_t *__restrict__ chunk_cone, double *__restrict__ chunk_sum, SocScale *__restrict__ scale,
             unsigned int *counter, RelaxArgs R, const Ctrl *ctrl)
{
    double ra = R.a, rb_ = R.b;
    if (RELAX && R.use_a12) {
        ra = ctrl->alpha12;
        rb_ = 1.0 - ra;
    }
    __shared__ double s_w[VBLOCK / 32];
    __shared__ bool s_last;
    const int ch = blockIdx.x;
    const SocCone C = cones[chunk_cone[ch]];
    const int64_t skip = C.rot ? 2 : 1;  // entries before the tail
    const int64_t k0 = (int64_t)(ch - C.chunk0) * SOC_CHUNK;  // offset inside the tail
    const int64_t tail = C.len - skip;
    const int64_t k1 = k0 + SOC_CHUNK < tail ? k0 + SOC_CHUNK : tail;
    double acc = 0.0;
    for (int64_t k = k0 + threadIdx.x; k < k1; k += VBLOCK) {
        const double w = relaxed_in<RELAX>(in, R, ra, rb_, C.head + skip + k);
        acc = fma(w, w, acc);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < VBLOCK / 32; k++) s += s_w[k];
        chunk_sum[ch] = s;
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int cidx = threadIdx.x; cidx < ncones; cidx += VBLOCK) {
        const SocCone K = cones[cidx];
        double s = 0.0;
        const volatile double *cs = chunk_sum;
        for (int k = 0; k < K.nchunk; k++) s += cs[K.chunk0 + k];
        if (RELAX) {  // plain SOCs only on this path (ConeSet::fusable): the classification needs the head entry
            const double head[1] = {relaxed_in<true>(in, R, ra, rb_, K.head)};
            SocCone K0 = K;
            K0.head = 0;
            scale[cidx] = soc_classify(K0, s, head);
        } else {
            scale[cidx] = soc_classify(K, s, in);
        }
    }
    if (threadIdx.x == 0) *counter = 0u;
}

__device__ __forceinline__ double cone_project(uint8_t op, double x, const double *__restrict__ proj, int64_t e,
                                               const int32_t *__restrict__ cone_of, const SocScale *__restrict__ scale,
                                               const double *in = nullptr, const double2 *__restrict__ box = nullptr,
                                               const SocCone *__restrict__ soc = nullptr)
{
    switch (op) {
    case OP_ZERO: return 0.0;
    case OP_COPY: return x;
    // NaN goes through, as in the reference (IndNonnegative: `x < 0 ? 0 : x`; Julia's max(NaN, 0) is NaN for tau and
    // kappa, cones.jl:138,141): after a CG breakdown on the indefinite KKT matrix the reference ends Indeterminate
    case OP_MAX0: return x < 0.0 ? 0.0 : x;
    case OP_MIN0: return x > 0.0 ? 0.0 : x;
    case OP_PRE: return proj[e];
    case OP_BOX: {  // IndBox: min(max(x, lo), hi)
        const double2 b = box[cone_of[e]];
        return x < b.x ? b.x : (x > b.y ? b.y : x);
    }
    default: break;
    }
    if (op >= OP_EXPP && op <= OP_EXPD_D) {
        const int pos = cone_of[e];
        const double *p0 = in + (e - pos);
        const double v[3] = {p0[0], p0[1], p0[2]};
        double y[3];
        if (op == OP_EXPP) {  // P_K(x)
            proj_exp_cone(v, y);
            return y[pos];
        }
        if (op == OP_EXPP_D || op == OP_EXPD) {  // x + P_K(-x)   (cones.jl:80-85 / IndExpDual by Moreau)
            const double nv[3] = {-v[0], -v[1], -v[2]};
            proj_exp_cone(nv, y);
            return add_(x, y[pos]);
        }
        // proxDual! of ExpDual: x + P_K*(-x),  P_K*(-x) = -x + P_K(x)
        proj_exp_cone(v, y);
        return add_(x, add_(-x, y[pos]));
    }
    const SocScale S = scale[cone_of[e]];
    if (op == OP_SOC_HEAD) return S.mode == 0 ? 0.0 : (S.mode == 1 ? x : mul_(S.rho, S.nx));
    if (op == OP_SOC_TAIL) return S.mode == 0 ? 0.0 : (S.mode == 1 ? x : mul_(S.rho, x));
    // dual: y = x + P(-x)
    if (op == OP_SOCD_HEAD) {
        const double pj = S.mode == 0 ? 0.0 : (S.mode == 1 ? -x : mul_(S.rho, S.nx));
        return add_(x, pj);
    }
    if (op == OP_SOCD_TAIL) {
        const double pj = S.mode == 0 ? 0.0 : (S.mode == 1 ? -x : mul_(S.rho, -x));
        return add_(x, pj);
    }
    // rotated SOC
    const bool dual = op >= OP_RSOCD_H0;
    const double sg = dual ? -1.0 : 1.0;
    double pj;
    if (op == OP_RSOC_TAIL || op == OP_RSOCD_TAIL) {
        pj = S.mode == 0 ? 0.0 : (S.mode == 1 ? sg * x : mul_(S.rho, sg * x));
    } else {
        const int64_t head = soc[cone_of[e]].head;
        double r1, r2;
        rsoc_rotate(sg * in[head], sg * in[head + 1], r1, r2);
        const double y1 = S.mode == 0 ? 0.0 : (S.mode == 1 ? r1 : mul_(S.rho, S.nx));
        const double y2 = S.mode == 0 ? 0.0 : (S.mode == 1 ? r2 : mul_(S.rho, r2));
        const bool first = (op == OP_RSOC_H0 || op == OP_RSOCD_H0);
        pj = first ? add_(mul_(RSOC_C, y1), mul_(RSOC_C, y2)) : sub_(mul_(RSOC_C, y1), mul_(RSOC_C, y2));
    }
    return dual ? add_(x, pj) : pj;
}

enum { EPI_NONE = 0, EPI_GAP = 1, EPI_GAPA = 2, EPI_FISTA = 3, EPI_DYKSTRA = 4, EPI_GAPP_PROJ = 5, EPI_LS = 6, EPI_LSW = 7 };

struct EpiArgs {
    double a2, om_a2;  // alpha2, 1 - alpha2
    double a, om_a;    // alpha, 1 - alpha
    double coef;       // FISTA (told-1)/t
    double betaA;      // GAPA beta
    double *tmp2;      // relaxed S2 output
    double *x;         // iterate (updated in place)
    double *aux1;      // FISTA: xold (out) ; Dykstra: q (in/out)
    double *aux2;      // FISTA: y (out)    ; Dykstra: y (in)
    double ls_alpha;   // EPI_LS / EPI_LSW: the step length tested
    int use_a12;       // EPI_GAPP_PROJ / EPI_LSW: take alpha2 from ctrl->alpha12 (GAPA under LineSearchWrapper)
    int coef_from_ctrl;  // EPI_FISTA on the graph path: coef = ctrl->fista_coef (set by k_iter_begin)
};

// proj = P_S2(in);  then the algorithm-specific epilogue:
//   EPI_GAP   (gap.jl:58,78)        tmp2 = a2*proj + (1-a2)*in ; x = a*tmp2 + (1-a)*x
//   EPI_GAPA  (gapa.jl:77,96-103)   same with a2 = alpha12, plus the angle estimate -> new alpha12
//   EPI_FISTA (fista.jl:39-46)      xold = x ; x = proj ; y = x + coef*(x - xold)
//   EPI_DYKSTRA (dykstra.jl:32-35)  x = proj ; q = in - x        (in = y + q)
//   EPI_GAPP_PROJ (gapproj.jl:61-62) tmp2 = a2*proj + (1-a2)*in ; x = tmp2
//   EPI_LS    (gapproj.jl:49-55)    ||proj - in|| ; last block keeps the strictly smallest
//   EPI_LSW   (wrappers/linesearch.jl:63-69)  t3 = a2*proj + (1-a2)*in ; ||x - t3|| ; strictly smallest
template <int EPI, bool RELAX = false>
__global__ void __launch_bounds__(VBLOCK)
k4_cone_apply(int64_t NP, const double *__restrict__ in, double *__restrict__ proj, const uint8_t *__restrict__ ops,
              const int32_t *__restrict__ cone_of, const SocScale *__restrict__ scale, const SocCone *__restrict__ soc,
              const double2 *__restrict__ box, EpiArgs E, Ctrl *ctrl, RedBuf rb, RelaxArgs R)
{
    double ra = R.a, rb_ = R.b;
    if (RELAX && R.use_a12) {
        ra = ctrl->alpha12;
        rb_ = 1.0 - ra;
    }
    double a2 = E.a2, om_a2 = E.om_a2;
    const double fcoef = (EPI == EPI_FISTA && E.coef_from_ctrl) ? ctrl->fista_coef : E.coef;
    if (EPI == EPI_GAPA || ((EPI == EPI_GAPP_PROJ || EPI == EPI_LSW) && E.use_a12)) {
        a2 = ctrl->alpha12;
        om_a2 = 1.0 - a2;
    }
    double q[3] = {0, 0, 0};
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        const double t1 = relaxed_in<RELAX>(in, R, ra, rb_, e);
        if (RELAX) R.tmp1[e] = t1;
        const double pj = cone_project(ops[e], t1, proj, e, cone_of, scale, in, box, soc);
        proj[e] = pj;
        if (EPI == EPI_GAP || EPI == EPI_GAPA) {
            const double t2 = add_(mul_(a2, pj), mul_(om_a2, t1));
            E.tmp2[e] = t2;
            const double xo = E.x[e];
            if (EPI == EPI_GAPA) {  // normedScalar(tmp2,tmp1,tmp1,x)  gapa.jl:36-47
                const double d1 = sub_(t2, t1), d2 = sub_(t1, xo);
                q[0] = fma(d1, d2, q[0]);
                q[1] = fma(d1, d1, q[1]);
                q[2] = fma(d2, d2, q[2]);
            }
            E.x[e] = add_(mul_(E.a, t2), mul_(E.om_a, xo));
        } else if (EPI == EPI_GAPP_PROJ) {
            const double t2 = add_(mul_(a2, pj), mul_(om_a2, t1));
            E.tmp2[e] = t2;
            E.x[e] = t2;
        } else if (EPI == EPI_FISTA) {
            const double xo = E.x[e];
            E.aux1[e] = xo;
            E.x[e] = pj;
            E.aux2[e] = add_(pj, mul_(fcoef, sub_(pj, xo)));
        } else if (EPI == EPI_DYKSTRA) {
            E.x[e] = pj;
            E.aux1[e] = sub_(t1, pj);
        } else if (EPI == EPI_LS) {
            const double d = sub_(pj, t1);
            q[0] = fma(d, d, q[0]);
        } else if (EPI == EPI_LSW) {
            const double t3 = add_(mul_(a2, pj), mul_(om_a2, t1));
            const double d = sub_(E.x[e], t3);  // normdiff(x, tmp3), linesearch.jl:64,77-85
            q[0] = fma(d, d, q[0]);
        }
    }
    if (EPI == EPI_GAPA) {
        double tot[3];
        if (grid_reduce<3, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
            double scl = fabs(tot[0]) / sqrt(tot[1] * tot[2]);  // normedScalar, gapa.jl:47
            if (isnan(scl)) scl = 0.0;                          // clamp keeps NaN, :97 maps it to 0
            else scl = scl < 0.0 ? 0.0 : (scl > 1.0 ? 1.0 : scl);  // clamp(scl, 0, 1)  :96
            const double s = sqrt(1.0 - scl * scl);
            const double aopt = 2.0 / (1.0 + s);
            ctrl->alpha12 = (1.0 - E.betaA) * aopt + E.betaA * 2.0;
        }
    } else if (EPI == EPI_LS || EPI == EPI_LSW) {
        double tot[3];
        if (grid_reduce<3, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
            const double nt = sqrt(tot[0]);
            if (nt < ctrl->ls_normbest) {  // gapproj.jl:52-55
                ctrl->ls_normbest = nt;
                ctrl->ls_alphabest = E.ls_alpha;
            }
        }
    }
}

// normbest = Inf ; alpha_best = -1.0 (gapproj.jl:44-45) or 1.0 (wrappers/linesearch.jl:55-56)
static __global__ void k_ls_begin(Ctrl *ctrl, double alpha_init)
{
    ctrl->ls_normbest = INFINITY;
    ctrl->ls_alphabest = alpha_init;
}

// =======================================================================================
// K6: status checks
// =======================================================================================
// HSDE (HSDEStatus.jl:27-71).  V = (A*z.x, A'*z.y), one right-hand side.
static __global__ void __launch_bounds__(VBLOCK)
k6_check_hsde(Lay L, MVView V, const double *__restrict__ z, const double *__restrict__ c,
              const double *__restrict__ b, double nb, double ncn, double eps, int64_t iter_i, int cgiter_host,
              Ctrl *ctrl, double *__restrict__ recs, int rec_cap, RedBuf rb)
{
    const int64_t LP = L.LP, oy = L.n_pad, ot = L.n_pad + L.m_pad;
    const double tau = z[ot], kap = z[LP + ot];
    double q[6] = {0, 0, 0, 0, 0, 0};  // P, D, ctx, bty, U, I
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < ot; e += (int64_t)gridDim.x * VBLOCK) {
        if (e < oy) {
            if (e < L.n) {
                const double aty = mv_atw(V, 0, e);
                const double x = z[e], r = z[LP + e], cj = c[e];
                const double dv = sub_(add_(aty / tau, cj), r / tau);  // :35
                q[1] = fma(dv, dv, q[1]);
                q[2] = fma(cj, x, q[2]);   // :36
                q[5] = fma(aty, aty, q[5]);  // :61
            }
        } else {
            const int64_t i = e - oy;
            if (i < L.m) {
                const double ax = mv_ax(V, 0, i);
                const double y = z[e], s = z[LP + e], bi = b[i];
                const double pv = sub_(add_(ax / tau, s / tau), bi);  // :34
                q[0] = fma(pv, pv, q[0]);
                q[3] = fma(bi, y, q[3]);   // :37
                const double uv = add_(ax, s);  // :59
                q[4] = fma(uv, uv, q[4]);
            }
        }
    }
    double tot[6];
    if (grid_reduce<6, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        const double p = sqrt(tot[0]) / fabs(1.0 + nb);
        const double d = sqrt(tot[1]) / fabs(1.0 + ncn);
        const double ctx = tot[2], bty = tot[3];
        const double g = fabs(ctx / tau + bty / tau) / (1.0 + fabs(ctx / tau) + fabs(bty / tau));  // :38
        int status = FOS_STATUS_CONTINUE;
        if (p <= eps * (1.0 + nb) && d <= eps * (1.0 + ncn) &&
            g <= eps * (1.0 + fabs(ctx / tau) + fabs(bty / tau)))  // :54
            status = FOS_STATUS_OPTIMAL;
        else if (sqrt(tot[4]) <= eps * (-ctx / ncn))  // :59
            status = FOS_STATUS_UNBOUNDED;
        else if (sqrt(tot[5]) <= eps * (-bty / nb))  // :61
            status = FOS_STATUS_INFEASIBLE;
        const int k = ctrl->nrec;
        if (k < rec_cap) {
            double *R = recs + (size_t)k * FOS_REC_LEN;
            R[0] = iter_i >= 0 ? (double)iter_i : (double)ctrl->cur_i;  // graph path: the device counts the iterations
            R[1] = p; R[2] = d; R[3] = g; R[4] = ctx; R[5] = bty; R[6] = kap; R[7] = tau;
            R[8] = cgiter_host >= 0 ? (double)cgiter_host : (double)ctrl->iter;
            R[9] = (double)status;
        }
        ctrl->nrec = k + 1;
        ctrl->status = status;
    }
}

// Feasibility form (FeasibilityStatus.jl:32-72): err = ||prev - z||, prev = z every iteration
static __global__ void __launch_bounds__(VBLOCK)
k6_check_feas(int64_t NP, const double *__restrict__ z, double *__restrict__ prev, int do_check, double eps,
              int64_t iter_i, Ctrl *ctrl, double *__restrict__ recs, int rec_cap, RedBuf rb)
{
    double q[1] = {0};
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        const double ze = z[e];
        const double d = sub_(prev[e], ze);
        q[0] = fma(d, d, q[0]);
        prev[e] = ze;
    }
    if (!do_check) return;
    double tot[1];
    if (grid_reduce<1, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        // prev starts as NaN (Feasibility.jl:79, begin_solve): the first error is NaN and never <= eps
        const double err = sqrt(tot[0]);
        const int status = (err <= eps) ? FOS_STATUS_OPTIMAL : FOS_STATUS_CONTINUE;
        const int k = ctrl->nrec;
        if (k < rec_cap) {
            double *R = recs + (size_t)k * FOS_REC_LEN;
            R[0] = (double)iter_i; R[1] = err;
            for (int t = 2; t < 8; t++) R[t] = 0.0;
            R[8] = (double)ctrl->iter;
            R[9] = (double)status;
        }
        ctrl->nrec = k + 1;
        ctrl->status = status;
    }
}

// =======================================================================================
// pack / unpack between the reference's contiguous vector and the padded device layout
// =======================================================================================
struct SegMap {
    int32_t nseg;
    int64_t len[6], src[6], dst[6];  // logical offset, padded offset
};
static __global__ void __launch_bounds__(VBLOCK)
k_pack(SegMap M, const double *__restrict__ logical, double *__restrict__ padded, int64_t NP)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < NP; e += (int64_t)gridDim.x * VBLOCK) {
        double v = 0.0;
        for (int s = 0; s < M.nseg; s++) {
            const int64_t k = e - M.dst[s];
            if (k >= 0 && k < M.len[s]) v = logical[M.src[s] + k];
        }
        padded[e] = v;
    }
}
static __global__ void __launch_bounds__(VBLOCK)
k_unpack(SegMap M, const double *__restrict__ padded, double *__restrict__ logical, int64_t N)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < N; e += (int64_t)gridDim.x * VBLOCK) {
        for (int s = 0; s < M.nseg; s++) {
            const int64_t k = e - M.src[s];
            if (k >= 0 && k < M.len[s]) logical[e] = padded[M.dst[s] + k];
        }
    }
}

// multi-GPU: fold the local partials into the exchange buffer [NV][n_pad] | [NV][m_pad].  Rows owned
// by other ranks (and all padding) are written as zeros on every pass: the all-reduce is in place, so
// the buffer still holds the previous pass's global sums.
template <int NV>
__global__ void __launch_bounds__(VBLOCK)
k1_finalize_local(MVView V, int64_t n, int64_t n_pad, int64_t m_local, int64_t row_begin, int64_t m_pad,
                  double *__restrict__ xbuf, const int32_t *skip_flag)
{
    if (skip_flag != nullptr && *skip_flag != 0) return;
    const int64_t total = n_pad + m_pad;
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < total; e += (int64_t)gridDim.x * VBLOCK) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
            if (e < n_pad) {
                xbuf[(size_t)v * n_pad + e] = e < n ? mv_atw(V, v, e) : 0.0;
            } else {
                const int64_t row = e - n_pad, lr = row - row_begin;
                xbuf[(size_t)NV * n_pad + (size_t)v * m_pad + row] = (lr >= 0 && lr < m_local) ? mv_ax(V, v, lr) : 0.0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K7, fused exchange over NVLink peer memory (one process per GPU, CUDA IPC): replaces
// k1_finalize_local + ncclAllReduce with ONE kernel, and is the first phase of the fused CG tail.
//
// Protocol ("self-validating 8-byte pushes"): every rank owns a receive area
//     recv[parity][source rank][vector 0..1][entry 0 .. n_pad+m_pad)
// that is filled with a SENTINEL bit pattern (all ones: a NaN no arithmetic produces).
//   push  a thread folds its entries of the local K1 partials -- column entries (A' partial sums) on every
//         rank, row entries (A X) on the rank that owns the row -- and stores each value straight into the
//         receive area of EVERY rank (its own included) with one 8-byte system-scope store per destination.
//         Remote stores are posted: nothing waits for them.
//   poll  the same thread then reads the same entries of its OWN receive area until they differ from the
//         sentinel, adds the sources in rank order (identical bits on every rank -> identical stop tests) and
//         puts the sentinel back.
// An aligned 8-byte store is indivisible, so a value is its own "ready" flag: there is no flag, no fence and no
// round trip on the critical path (the previous flag protocol spent 3 system-scope fences, ~4 us each, and a
// remote-load round trip per exchange: profiles/r2_tail_trace.md).  Areas are double buffered by exchange
// parity: a rank writes parity p again two exchanges later, after it has consumed the intermediate exchange, which
// the receiver only feeds after its own kernel of exchange p -- and with it every reset -- has completed.
// ---------------------------------------------------------------------------------------
constexpr int P2P_MAX_RANKS = 8;   // one NVSwitch domain
constexpr int P2P_MAX_BLOCKS = 512;  // grid limit of k_cg_tail_hsde<true>
constexpr unsigned long long P2P_SENTINEL = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long P2P_CANONICAL_NAN = 0x7FF8000000000000ull;
struct P2PHeader {  // first bytes of every rank's region
    int64_t row_begin, m_local;
    int64_t reserved[14];
};
struct P2PView {
    unsigned char *peer[P2P_MAX_RANKS];  // base of every rank's region (own region included), mapped here
    int64_t row_begin[P2P_MAX_RANKS], m_local[P2P_MAX_RANKS];
    int32_t nranks, rank;
    int64_t recv_off;        // byte offset of the receive area inside a region
    int64_t E;               // entries per vector: n_pad + m_pad
    unsigned int *epoch;     // local: exchanges completed so far
    unsigned int *tickets;   // local: [0] exit ticket of k1_exchange_p2p
    unsigned int *error;     // local: set when a peer did not show up within P2P_TIMEOUT_CYCLES
    int32_t tail_flag_mode;  // unused (kept so that the option "tail_flags" stays accepted)
};

// A peer that died (or never launched the matching kernel) must not hang this GPU: the polls give up after
// ~30 s of SM clock, raise the error flag (the host turns it into FOS_ERR_COMM at the next synchronisation) and
// let the kernel finish with whatever it has.
constexpr long long P2P_TIMEOUT_CYCLES = 60000000000LL;

__device__ __forceinline__ unsigned long long *p2p_recv(const P2PView &X, int dst_rank, int par, int src_rank, int v,
                                                        int64_t e)
{
    return reinterpret_cast<unsigned long long *>(X.peer[dst_rank] + X.recv_off) +
           ((((size_t)par * X.nranks + src_rank) * 2 + v) * (size_t)X.E + (size_t)e);
}
// store `val` as entry e / vector v of this rank into the receive area of every rank
__device__ __forceinline__ void p2p_push(const P2PView &X, int par, int v, int64_t e, double val)
{
    unsigned long long bits = (unsigned long long)__double_as_longlong(val);
    if (bits == P2P_SENTINEL) bits = P2P_CANONICAL_NAN;
    for (int r = 0; r < X.nranks; r++)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p2p_recv(X, r, par, X.rank, v, e)), "l"(bits) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Waits for NV values of entry e from every source in [r0, r1) (all first loads in flight together), returns
// them and re-arms the slots.  out[k][v], k = source - r0.
template <int NV>
__device__ __forceinline__ void p2p_poll(const P2PView &X, int par, int64_t e, int r0, int r1,
                                         double (&out)[P2P_MAX_RANKS][NV])
{
    unsigned long long bits[P2P_MAX_RANKS][NV];
#pragma unroll
    for (int k = 0; k < P2P_MAX_RANKS; k++)
        if (r0 + k < r1)
#pragma unroll
            for (int v = 0; v < NV; v++) bits[k][v] = ld_sys_u64(p2p_recv(X, X.rank, par, r0 + k, v, e));
    long long t0 = 0;
    bool timing = false;
#pragma unroll
    for (int k = 0; k < P2P_MAX_RANKS; k++)
        if (r0 + k < r1)
#pragma unroll
            for (int v = 0; v < NV; v++) {
                unsigned long long *slot = p2p_recv(X, X.rank, par, r0 + k, v, e);
                while (bits[k][v] == P2P_SENTINEL) {
                    if (!timing) {
                        timing = true;
                        t0 = clock64();
                    } else if (clock64() - t0 > P2P_TIMEOUT_CYCLES) {
                        atomicExch(X.error, 1u);
                        bits[k][v] = P2P_CANONICAL_NAN;
                        break;
                    }
                    bits[k][v] = ld_sys_u64(slot);
                }
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(P2P_SENTINEL) : "memory");
                out[k][v] = __longlong_as_double((long long)bits[k][v]);
            }
}
// Sum over ranks (rank order) of the NV values of column entry e.
template <int NV>
__device__ __forceinline__ void p2p_poll_sum(const P2PView &X, int par, int64_t e, double (&acc)[NV])
{
    double t[P2P_MAX_RANKS][NV];
    p2p_poll<NV>(X, par, e, 0, X.nranks, t);
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
#pragma unroll
    for (int k = 0; k < P2P_MAX_RANKS; k++)
        if (k < X.nranks)
#pragma unroll
            for (int v = 0; v < NV; v++) acc[v] += t[k][v];
}
__device__ __forceinline__ int p2p_row_owner(const P2PView &X, int64_t row)
{
    int owner = -1;
    for (int r = 0; r < X.nranks; r++)
        if (row >= X.row_begin[r] && row < X.row_begin[r] + X.m_local[r]) owner = r;
    return owner;
}
// the NV values of row entry `row` from its owner
template <int NV>
__device__ __forceinline__ void p2p_poll_row(const P2PView &X, int par, int64_t n_pad, int64_t row, double (&val)[NV])
{
    const int owner = p2p_row_owner(X, row);
    if (owner < 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) val[v] = 0.0;
        return;
    }
    double t[P2P_MAX_RANKS][NV];
    p2p_poll<NV>(X, par, n_pad + row, owner, owner + 1, t);
#pragma unroll
    for (int v = 0; v < NV; v++) val[v] = t[0][v];
}
// push phase of one exchange: the thread's entries e0, e0 + stride, ... of [0, n_pad + m_pad)
template <int NV>
__device__ __forceinline__ void p2p_push_all(const MVView &V, const P2PView &X, int par, int64_t n, int64_t n_pad,
                                             int64_t m_pad, int64_t e0, int64_t stride)
{
    const int64_t rb = X.row_begin[X.rank], ml = X.m_local[X.rank];
    for (int64_t e = e0; e < n_pad + m_pad; e += stride) {
        if (e < n_pad) {
            if (e < n) {
                if (NV == 2) {
                    double w0, w1;
                    mv_atw2(V, e, w0, w1);
                    p2p_push(X, par, 0, e, w0);
                    p2p_push(X, par, 1, e, w1);
                } else {
                    p2p_push(X, par, 0, e, mv_atw(V, 0, e));
                }
            }
        } else {
            const int64_t lr = e - n_pad - rb;
            if (lr >= 0 && lr < ml) {
                if (NV == 2) {
                    double a0, a1;
                    mv_ax2(V, lr, a0, a1);
                    p2p_push(X, par, 0, e, a0);
                    p2p_push(X, par, 1, e, a1);
                } else {
                    p2p_push(X, par, 0, e, mv_ax(V, 0, lr));
                }
            }
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(VBLOCK)
k1_exchange_p2p(MVView V, P2PView X, int64_t n, int64_t n_pad, int64_t m_pad, double *__restrict__ xbuf,
                const int32_t *skip_flag)
{
    if (skip_flag != nullptr && *skip_flag != 0) return;
    const unsigned int epoch = *X.epoch + 1u;
    const int par = (int)(epoch & 1u);
    const int64_t total = n_pad + m_pad;
    const int64_t e0 = (int64_t)blockIdx.x * VBLOCK + threadIdx.x, stride = (int64_t)gridDim.x * VBLOCK;
    p2p_push_all<NV>(V, X, par, n, n_pad, m_pad, e0, stride);
    for (int64_t e = e0; e < total; e += stride) {
        double val[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) val[v] = 0.0;
        if (e < n_pad) {
            if (e < n) p2p_poll_sum<NV>(X, par, e, val);
#pragma unroll
            for (int v = 0; v < NV; v++) xbuf[(size_t)v * n_pad + e] = val[v];
        } else {
            const int64_t row = e - n_pad;
            p2p_poll_row<NV>(X, par, n_pad, row, val);
#pragma unroll
            for (int v = 0; v < NV; v++) xbuf[(size_t)NV * n_pad + (size_t)v * m_pad + row] = val[v];
        }
    }
    // ---- exit: the last block advances the epoch and re-arms the ticket ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(&X.tickets[0], 1u);
        if (t == gridDim.x - 1) {
            X.tickets[0] = 0u;
            *X.epoch = epoch;
            __threadfence();
        }
    }
}

// Initial CG residual on row shards: k1_exchange_p2p + k2_kkt_hsde<K2_RESID> in ONE kernel (graph path and
// Handle::cg_solve with the peer exchange).  r = rhs - [I Q'; Q -I] x0, p = r, rn = <r,r>, iter = 1
// (conjugategradients.jl:32-36).  Blocks wait for their peers only, never for each other: an ordinary launch whose
// grid is resident (<= 2 blocks per SM) is enough.
static __global__ void __launch_bounds__(VBLOCK)
k2_resid_hsde_p2p(Lay L, MVView V, P2PView X, const double *__restrict__ in, const double *__restrict__ c,
                  const double *__restrict__ b, const double *__restrict__ rhs, double *__restrict__ r,
                  double *__restrict__ p, Ctrl *ctrl, RedBuf rb)
{
    const int64_t LP = L.LP, oy = L.n_pad, ot = L.n_pad + L.m_pad;
    const int64_t stride = (int64_t)gridDim.x * VBLOCK, e0 = (int64_t)blockIdx.x * VBLOCK + threadIdx.x;
    const unsigned int epoch = *X.epoch + 1u;
    const int par = (int)(epoch & 1u);
    p2p_push_all<2>(V, X, par, L.n, L.n_pad, L.m_pad, e0, stride);
    const double tau1 = in[ot], tau2 = in[LP + ot];
    double q[5] = {0, 0, 0, 0, 0};
    for (int64_t e = e0; e < ot; e += stride) {
        double o1 = 0.0, o2 = 0.0;
        const double i1 = in[e], i2 = in[LP + e];
        if (e < oy) {
            if (e < L.n) {
                double w[2];
                p2p_poll_sum<2>(X, par, e, w);
                const double cj = c[e];
                const double q1 = add_(w[0], mul_(tau1, cj));
                const double q2 = add_(w[1], mul_(tau2, cj));
                o1 = add_(-q2, i1);
                o2 = sub_(q1, i2);
                q[0] = fma(cj, i1, q[0]);
                q[2] = fma(cj, i2, q[2]);
            }
        } else {
            const int64_t i = e - oy;
            if (i < L.m) {
                double a[2];
                p2p_poll_row<2>(X, par, L.n_pad, i, a);
                const double bi = b[i];
                const double q1 = -sub_(a[0], mul_(tau1, bi));
                const double q2 = -sub_(a[1], mul_(tau2, bi));
                o1 = add_(-q2, i1);
                o2 = sub_(q1, i2);
                q[1] = fma(bi, i1, q[1]);
                q[3] = fma(bi, i2, q[3]);
            }
        }
        const double r1 = sub_(rhs[e], o1), r2 = sub_(rhs[LP + e], o2);
        r[e] = r1;
        r[LP + e] = r2;
        p[e] = r1;
        p[LP + e] = r2;
        q[4] = fma(r1, r1, q[4]);
        q[4] = fma(r2, r2, q[4]);
    }
    double tot[5];
    if (grid_reduce<5, VBLOCK>(q, tot, rb) && threadIdx.x == 0) {
        const double q1t = sub_(-tot[0], tot[1]);
        const double q2t = sub_(-tot[2], tot[3]);
        const double o1 = add_(-q2t, tau1);
        const double o2 = sub_(q1t, tau2);
        const double r1 = sub_(rhs[ot], o1), r2 = sub_(rhs[LP + ot], o2);
        r[ot] = r1;
        r[LP + ot] = r2;
        p[ot] = r1;
        p[LP + ot] = r2;
        ctrl->rn = tot[4] + r1 * r1 + r2 * r2;
        ctrl->iter = 1;
        *X.epoch = epoch;  // every block has finished its polls before the last one arrives here
    }
}

// =======================================================================================
// Fused CG tail (HSDE form): everything of one CG iteration after the pass over A, in ONE
// cooperative kernel --  [peer exchange]  ->  Ap = KKT p, <Ap,p>  ->  alpha  ->  x += alpha p,
// r -= alpha Ap, ||r||  ->  stop test, beta  ->  p = beta p + r  (conjugategradients.jl:38-51 on
// top of affinepluslinear.jl:37-49 / HSDEAffine.jl:41-65).  Replaces k1_exchange_p2p + k2_kkt_hsde<K2_AP>
// + k3_cg_update + k3_cg_dir: one launch instead of four, the vectors make one trip through the
// SMs instead of three, and the two dot products are grid reductions (fixed tree, bitwise
// reproducible) separated by grid barriers instead of kernel boundaries.
// Every thread owns the same entries e, LP+e in all phases, so the only cross-thread traffic is
// the reductions.  With P2P the gathered A X / A' W entries are read straight from the peers'
// exchange slots (see k1_exchange_p2p for the slot protocol, shared with this kernel).
// =======================================================================================
struct GridBar {
    unsigned int *count;  // arrivals since the handle was created (monotonic)
    unsigned int *flag;   // last completed barrier target, on its own cache line
    unsigned int *base;   // value of *count when the current launch started
    unsigned int *exit_ticket;
    double *part;         // [2][gridDim.x][8] block partials of the two reductions
    unsigned long long *trace;  // optional ("tail_trace"): [3 blocks][16] summed SM cycles per phase, [..][15] = launches
};
// phase timing of the fused CG tail (debug option "tail_trace"): thread 0 of the first, middle and last block adds
// the SM cycles since its previous mark to trace[block slot][phase]
#define FOS_TAIL_MARK(PH)                                                                \
    do {                                                                                 \
        if (tr_slot >= 0) {                                                              \
            const long long tnow_ = clock64();                                           \
            atomicAdd(gb.trace + tr_slot * 16 + (PH), (unsigned long long)(tnow_ - tr_last)); \
            tr_last = tnow_;                                                             \
        }                                                                                \
    } while (0)
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Arrivals are counted with one atomic per block on `count`; the last arriver publishes the target in
// `flag`, which lives on its own cache line: the waiting blocks poll that line (plain reads, with a
// short sleep) and never compete with the arriving atomics for the counter's line.
__device__ __forceinline__ void grid_barrier(const GridBar &gb, unsigned int &k, unsigned int base)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        k++;
        const unsigned int target = base + k * gridDim.x;
        __threadfence();
        const unsigned int old = atomicAdd(gb.count, 1u);
        if (old + 1u == target) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(gb.flag), "r"(target) : "memory");
        } else {
            while ((int)(ld_relaxed_gpu_u32(gb.flag) - target) < 0) __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}
// block partials -> part[blockIdx.x][0..NQ), grid barrier, every block adds all partials in block order
template <int NQ>
__device__ __forceinline__ void grid_allreduce(double (&q)[NQ], double *part, const GridBar &gb, unsigned int &k,
                                               unsigned int base)
{
    __shared__ double s_w[8][VBLOCK / 32];
    __shared__ double s_tot[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NQ; i++) {
        const double v = warp_sum(q[i]);
        if (lane == 0) s_w[i][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        double s = 0.0;
        for (int j = 0; j < VBLOCK / 32; j++) s += s_w[threadIdx.x][j];
        part[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
    }
    grid_barrier(gb, k, base);
    // fixed tree: thread t adds blocks t, t+VBLOCK, ...; warp shuffle tree; 8 warp sums in order
    {
        double acc[NQ];
#pragma unroll
        for (int i = 0; i < NQ; i++) acc[i] = 0.0;
        for (unsigned int bidx = threadIdx.x; bidx < gridDim.x; bidx += VBLOCK) {
            double t[NQ];
#pragma unroll
            for (int i = 0; i < NQ; i++) t[i] = __ldcg(part + (size_t)bidx * 8 + i);  // all in flight together
#pragma unroll
            for (int i = 0; i < NQ; i++) acc[i] += t[i];
        }
#pragma unroll
        for (int i = 0; i < NQ; i++) {
            const double v = warp_sum(acc[i]);
            if (lane == 0) s_w[i][w] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        double s = 0.0;
        for (int j = 0; j < VBLOCK / 32; j++) s += s_w[threadIdx.x][j];
        s_tot[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NQ; i++) q[i] = s_tot[i];
    __syncthreads();
}

template <bool P2P>
__global__ void __launch_bounds__(VBLOCK)
k_cg_tail_hsde(Lay L, MVView V, P2PView X, const double *__restrict__ c, const double *__restrict__ b,
               double *__restrict__ sol, double *__restrict__ r, double *__restrict__ p, double *__restrict__ Ap,
               Ctrl *ctrl, GridBar gb, unsigned long long cond)
{
    // `cond`: handle of the CUDA-graph WHILE node whose body this launch belongs to (0 outside a graph): the loop
    // goes round again until the stop test of conjugategradients.jl:42 fires
    if (cg_skip(ctrl)) return;
    int tr_slot = -1;
    long long tr_last = 0;
    if (gb.trace != nullptr && threadIdx.x == 0) {
        tr_slot = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : (blockIdx.x == gridDim.x - 1 ? 2 : -1));
        tr_last = clock64();
    }
    const unsigned int base = *gb.base;
    unsigned int nbar = 0;
    const int64_t LP = L.LP, oy = L.n_pad, ot = L.n_pad + L.m_pad;
    const int64_t stride = (int64_t)gridDim.x * VBLOCK, e0 = (int64_t)blockIdx.x * VBLOCK + threadIdx.x;
    const double rn = ctrl->rn, tol = ctrl->tol;
    const int iter = ctrl->iter, max_iters = ctrl->max_iters;
    unsigned int epoch = 0;
    int par = 0;
    if (P2P) {
        // ---- exchange, push phase: fold the local partials, store them into every rank's receive area ----
        epoch = *X.epoch + 1u;
        par = (int)(epoch & 1u);
        p2p_push_all<2>(V, X, par, L.n, L.n_pad, L.m_pad, e0, stride);
        FOS_TAIL_MARK(0);  // fold + push
    }
    // ---- Ap = [I Q'; Q -I] p and the dot products (k2_kkt_hsde<K2_AP>) ----
    const double tau1 = p[ot], tau2 = p[LP + ot];
    double q[5] = {0, 0, 0, 0, 0};
    for (int64_t e = e0; e < ot; e += stride) {
        double o1 = 0.0, o2 = 0.0;
        const double i1 = p[e], i2 = p[LP + e];
        if (e < oy) {
            if (e < L.n) {
                double w0, w1;
                if (P2P) {
                    double acc[2];
                    p2p_poll_sum<2>(X, par, e, acc);
                    w0 = acc[0];
                    w1 = acc[1];
                } else {
                    mv_atw2(V, e, w0, w1);
                }
                const double cj = c[e];
                const double q1 = add_(w0, mul_(tau1, cj));
                const double q2 = add_(w1, mul_(tau2, cj));
                o1 = add_(-q2, i1);
                o2 = sub_(q1, i2);
                q[0] = fma(cj, i1, q[0]);
                q[2] = fma(cj, i2, q[2]);
            }
        } else {
            const int64_t i = e - oy;
            if (i < L.m) {
                double a0, a1;
                if (P2P) {
                    double av[2];
                    p2p_poll_row<2>(X, par, L.n_pad, i, av);
                    a0 = av[0];
                    a1 = av[1];
                } else {
                    mv_ax2(V, i, a0, a1);
                }
                const double bi = b[i];
                const double q1 = -sub_(a0, mul_(tau1, bi));
                const double q2 = -sub_(a1, mul_(tau2, bi));
                o1 = add_(-q2, i1);
                o2 = sub_(q1, i2);
                q[1] = fma(bi, i1, q[1]);
                q[3] = fma(bi, i2, q[3]);
            }
        }
        Ap[e] = o1;
        Ap[LP + e] = o2;
        q[4] = fma(o1, i1, q[4]);
        q[4] = fma(o2, i2, q[4]);
    }
    FOS_TAIL_MARK(3);  // gather / fold + Ap
    grid_allreduce<5>(q, gb.part, gb, nbar, base);
    FOS_TAIL_MARK(4);  // first grid all-reduce
    const double q1t = sub_(-q[0], q[1]);  // HSDEAffine.jl:57
    const double q2t = sub_(-q[2], q[3]);
    const double o1t = add_(-q2t, tau1);
    const double o2t = sub_(q1t, tau2);
    const double pAp = q[4] + o1t * tau1 + o2t * tau2;
    const double alpha = rn / pAp;  // cg :39
    // ---- x += alpha p ; r -= alpha Ap ; ||r||  (k3_cg_update) ----
    // r_new and p of the first TAIL_EPT entries of every thread stay in registers for the direction update
    // (one global round trip less on the critical path); longer vectors fall back to memory
    constexpr int TAIL_EPT = 3;
    double keep_r[TAIL_EPT][2], keep_p[TAIL_EPT][2];
    double q2r[1] = {0.0};
    {
        int kx = 0;
        for (int64_t e = e0; e < LP; e += stride, kx++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int64_t g = e + h * LP;
                double ape = Ap[g];
                if (e == ot) ape = h == 0 ? o1t : o2t;
                const double pe = p[g];
                sol[g] = add_(sol[g], mul_(alpha, pe));
                const double re = sub_(r[g], mul_(alpha, ape));
                r[g] = re;
                q2r[0] = fma(re, re, q2r[0]);
#pragma unroll
                for (int t = 0; t < TAIL_EPT; t++)
                    if (t == kx) {
                        keep_r[t][h] = re;
                        keep_p[t][h] = pe;
                    }
            }
        }
    }
    FOS_TAIL_MARK(5);  // x, r update
    grid_allreduce<1>(q2r, gb.part + (size_t)gridDim.x * 8, gb, nbar, base);
    FOS_TAIL_MARK(6);  // second grid all-reduce
    const double rr = q2r[0];
    const double rnorm = sqrt(rr);
    const bool stop = rnorm <= tol || iter >= max_iters;  // cg :42
    const double beta = rr / rn;
    if (!stop) {
        int kx = 0;
        for (int64_t e = e0; e < LP; e += stride, kx++) {  // p = beta p + r  (:49-50)
            double pe0, pe1, re0, re1;
            if (kx < TAIL_EPT) {
                pe0 = pe1 = re0 = re1 = 0.0;
#pragma unroll
                for (int t = 0; t < TAIL_EPT; t++)
                    if (t == kx) {
                        pe0 = keep_p[t][0];
                        pe1 = keep_p[t][1];
                        re0 = keep_r[t][0];
                        re1 = keep_r[t][1];
                    }
            } else {
                pe0 = p[e];
                pe1 = p[LP + e];
                re0 = r[e];
                re1 = r[LP + e];
            }
            p[e] = add_(mul_(beta, pe0), re0);
            p[LP + e] = add_(mul_(beta, pe1), re1);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Ap[ot] = o1t;
        Ap[LP + ot] = o2t;
        ctrl->alpha = alpha;
        ctrl->rnorm = rnorm;
        if (stop) {
            ctrl->done = 1;
            ctrl->total_cg_dev += iter;
            if (iter >= max_iters) ctrl->warn_maxit = 1;  // :53
        } else {
            ctrl->rn = rr;
            ctrl->beta = beta;
            ctrl->iter = iter + 1;
        }
        if (cond != 0ull) cudaGraphSetConditional((cudaGraphConditionalHandle)cond, stop ? 0u : 1u);
    }
    FOS_TAIL_MARK(7);  // direction update
    if (tr_slot >= 0) atomicAdd(gb.trace + tr_slot * 16 + 15, 1ull);
    // ---- exit: the last block re-arms the barrier base (and the exchange epoch / tickets) ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(gb.exit_ticket, 1u) == gridDim.x - 1) {
            *gb.exit_ticket = 0u;
            *gb.base = base + nbar * gridDim.x;
            if (P2P) *X.epoch = epoch;
            __threadfence();
        }
    }
}
#endif  // __CUDACC__

}  // namespace fos
