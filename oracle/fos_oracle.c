/*
 * fos_oracle.c -- CPU restatement of the FirstOrderSolvers.jl per-iteration hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (firstordersolvers.jl_b200/csrc) never links, calls or falls back to anything here.
 *
 * PARITY STATUS: "parity unpinned" for the seeded goldens of the reference
 * (test/testDRandGAPA.jl:12,15 and test/testprint.jl:19 need Julia's RNG and Convex.jl's
 * lowering, neither of which exists in this environment -- no julia binary).  What IS
 * pinned: the literal 2x2 PSD known-answer of test/testPSD.jl:3-4 and every property
 * test of test/conjugateGradient.jl, test/HSDEAffine.jl, test/affinepluslinear.jl and
 * test/testfeasibility.jl, re-expressed in tests/test_oracle_*.py against dense solves.
 * The cone arithmetic lives in ProximalOperators.jl (un-vendored, un-pinned dependency,
 * Project.toml:10); its published algorithm is restated in the cone section below.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src unless stated).  Arithmetic is FP64, compiled with
 * -ffp-contract=off so that broadcast expressions round like Julia's (no fused
 * multiply-add), and sparse products walk the CSC structure in the same order as
 * Julia's SparseArrays.mul! (column scatter for A*x, per-column dot for A'*y).
 *
 * Single-threaded by construction: the reference's mat-vecs and broadcasts are serial
 * (SURVEY.md F6).
 *
 * Build variants of this one file (oracle/Makefile):
 *   libfos_oracle.so      the restatement proper: FP64 throughout, sequential sums in the reference's order.
 *   libfos_oracle_hp.so   -DFOS_ORACLE_HP: every REDUCTION (dot, norm, the sums of a sparse product) is
 *                         accumulated in long double (64-bit mantissa) and rounded to FP64 once; element-wise
 *                         operations round exactly like the restatement.  This is the "exact" yardstick of the
 *                         parity tests: the distance of any FP64 implementation from it is that implementation's
 *                         own accumulated summation error (tests/test_gpu_exact.py).
 *   libfos_oracle_mt.so   -DFOS_ORACLE_MT -fopenmp: the sparse products and the CG sweeps are spread over the
 *                         host threads (bench.py --impl reference: "with all the host threads it can use").
 *                         Same algorithm, different summation order; used for TIMING only, never as a checker.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FOSOR_API __attribute__((visibility("default")))

#ifdef FOS_ORACLE_HP
typedef long double acc_t; /* accumulator of every reduction */
#else
typedef double acc_t;
#endif
#ifdef FOS_ORACLE_MT
#include <omp.h>
#define PFOR _Pragma("omp parallel for schedule(static)")
#else
#define PFOR
#endif

/* ------------------------------------------------------------------------------------
 * cone type codes (cones.jl:4-14, conemap)
 * ---------------------------------------------------------------------------------- */
enum {
    CONE_FREE = 0,
    CONE_ZERO = 1,
    CONE_NONNEG = 2,
    CONE_NONPOS = 3,
    CONE_SOC = 4,
    CONE_SOCROT = 5, /* IndRotatedSOC -- parity unpinned: no reference test touches it */
    CONE_SDP = 6,
    CONE_EXPPRIMAL = 7, /* IndExpPrimal -- parity unpinned, see prox_exp */
    CONE_EXPDUAL = 8    /* IndExpDual */
};

/* algorithm codes (solvers/{gap,gapa,fista,dykstra,gapproj}.jl) */
enum { ALG_GAP = 0, ALG_GAPA = 1, ALG_FISTA = 2, ALG_DYKSTRA = 3, ALG_GAPP = 4 };

/* status codes (HSDEStatus.jl:53-63) */
enum { ST_CONTINUE = 0, ST_OPTIMAL = 1, ST_UNBOUNDED = 2, ST_INFEASIBLE = 3, ST_INDETERMINATE = 4 };

/* ------------------------------------------------------------------------------------
 * small dense helpers: LinearAlgebra.dot / norm and broadcasts
 * ---------------------------------------------------------------------------------- */
static double vdot(const double *a, const double *b, int64_t n)
{
#ifdef FOS_ORACLE_MT
    double sm = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sm)
    for (int64_t i = 0; i < n; i++) sm += a[i] * b[i];
    return sm;
#else
    acc_t s = 0.0;
    for (int64_t i = 0; i < n; i++) s += (acc_t)a[i] * (acc_t)b[i];
    return (double)s;
#endif
}
static double vnorm(const double *a, int64_t n) { return sqrt(vdot(a, a, n)); }

/* ------------------------------------------------------------------------------------
 * CSC matrix = SparseMatrixCSC{Float64,Int64} (types.jl:35), stored 0-based here
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int64_t m, n;
    int64_t *colptr; /* n+1 */
    int64_t *rowval; /* nnz */
    double *nzval;   /* nnz */
    int borrowed;    /* arrays belong to the caller (fosor_create_conic_borrowed: no 12.8 GB copy at config 2) */
} csc_t;

static csc_t *csc_new(int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval,
                      const double *nzval, int64_t index_base)
{
    csc_t *A = (csc_t *)calloc(1, sizeof(csc_t));
    int64_t nnz = colptr[n] - index_base;
    A->m = m;
    A->n = n;
    A->colptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    A->rowval = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    A->nzval = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    for (int64_t j = 0; j <= n; j++) A->colptr[j] = colptr[j] - index_base;
    for (int64_t k = 0; k < nnz; k++) {
        A->rowval[k] = rowval[k] - index_base;
        A->nzval[k] = nzval[k];
    }
    return A;
}
static csc_t *csc_borrow(int64_t m, int64_t n, int64_t *colptr, int64_t *rowval, double *nzval)
{
    csc_t *A = (csc_t *)calloc(1, sizeof(csc_t));
    A->m = m;
    A->n = n;
    A->colptr = colptr; /* 0-based */
    A->rowval = rowval;
    A->nzval = nzval;
    A->borrowed = 1;
    return A;
}
static void csc_free(csc_t *A)
{
    if (!A) return;
    if (!A->borrowed) {
        free(A->colptr);
        free(A->rowval);
        free(A->nzval);
    }
    free(A);
}

/* y = A*x, Julia SparseArrays.mul!(y, A, x): zero y, then scatter column by column. */
static void csc_mul(double *y, const csc_t *A, const double *x)
{
#if defined(FOS_ORACLE_MT)
    /* column ranges per thread, private result vectors, added in thread order */
    const int64_t m = A->m, n = A->n;
    int nt = omp_get_max_threads();
    double *priv = (double *)calloc((size_t)nt * (size_t)(m > 0 ? m : 1), sizeof(double));
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        double *yt = priv + (size_t)t * (size_t)m;
        const int64_t j0 = n * t / nt, j1 = n * (t + 1) / nt;
        for (int64_t j = j0; j < j1; j++) {
            const double xj = x[j];
            for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; k++) yt[A->rowval[k]] += A->nzval[k] * xj;
        }
#pragma omp barrier
#pragma omp for schedule(static)
        for (int64_t i = 0; i < m; i++) {
            double sm = 0.0;
            for (int q = 0; q < nt; q++) sm += priv[(size_t)q * (size_t)m + (size_t)i];
            y[i] = sm;
        }
    }
    free(priv);
#elif defined(FOS_ORACLE_HP)
    acc_t *acc = (acc_t *)calloc((size_t)(A->m > 0 ? A->m : 1), sizeof(acc_t));
    for (int64_t j = 0; j < A->n; j++) {
        acc_t xj = x[j];
        for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; k++) acc[A->rowval[k]] += (acc_t)A->nzval[k] * xj;
    }
    for (int64_t i = 0; i < A->m; i++) y[i] = (double)acc[i];
    free(acc);
#else
    for (int64_t i = 0; i < A->m; i++) y[i] = 0.0;
    for (int64_t j = 0; j < A->n; j++) {
        double xj = x[j];
        for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; k++) y[A->rowval[k]] += A->nzval[k] * xj;
    }
#endif
}
/* y = A'*x, Julia SparseArrays.mul!(y, transpose(A), x): one dot product per column. */
static void csc_mul_t(double *y, const csc_t *A, const double *x)
{
    PFOR
    for (int64_t j = 0; j < A->n; j++) {
        acc_t t = 0.0;
        for (int64_t k = A->colptr[j]; k < A->colptr[j + 1]; k++) t += (acc_t)A->nzval[k] * (acc_t)x[A->rowval[k]];
        y[j] = (double)t;
    }
}

/* ------------------------------------------------------------------------------------
 * Linear operator used inside KKTMatrix / AffinePlusLinear: either a plain matrix or
 * the HSDE skew matrix Q (problemforms/HSDE/HSDEAffine.jl:2-20)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int kind;        /* 0 = plain CSC matrix, 1 = HSDEMatrixQ(A,b,c) */
    const csc_t *A;
    const double *b; /* length A->m (kind 1) */
    const double *c; /* length A->n (kind 1) */
    int64_t am, an;  /* operator size */
} linop_t;

/* HSDEMatrixQ mul!  (HSDEAffine.jl:41-59):  Y = Q*B,  Q = [0 A' c; -A 0 b; -c' -b' 0] */
static void hsdeq_mul(double *Y, const linop_t *Q, const double *B)
{
    const csc_t *A = Q->A;
    int64_t n = A->n, m = A->m;
    double *y1 = Y, *y2 = Y + n;
    const double *b1 = B, *b2 = B + n;
    double b3 = B[n + m];
    csc_mul_t(y1, A, b2);                                      /* :51 */
    csc_mul(y2, A, b1);                                        /* :52 */
    PFOR
    for (int64_t j = 0; j < n; j++) y1[j] += b3 * Q->c[j];     /* :54 */
    PFOR
    for (int64_t i = 0; i < m; i++) y2[i] -= b3 * Q->b[i];     /* :55 */
    PFOR
    for (int64_t i = 0; i < m; i++) y2[i] = -y2[i];            /* :56 */
    Y[n + m] = -vdot(Q->c, b1, n) - vdot(Q->b, b2, m);         /* :57 */
}

/* y = Op*x */
static void linop_mul(double *y, const linop_t *Op, const double *x)
{
    if (Op->kind == 0) csc_mul(y, Op->A, x);
    else hsdeq_mul(y, Op, x);
}
/* y = Op'*x ; for Q: mul! then negate (HSDEAffine.jl:61-65) */
static void linop_mul_t(double *y, const linop_t *Op, const double *x)
{
    if (Op->kind == 0) csc_mul_t(y, Op->A, x);
    else {
        hsdeq_mul(y, Op, x);
        PFOR
        for (int64_t i = 0; i < Op->an; i++) y[i] = -y[i];
    }
}

/* KKTMatrix mul! (utilities/affinepluslinear.jl:37-49):  y = [I A'; A -I] x */
static void kkt_mul(double *y, const linop_t *Op, const double *x)
{
    int64_t an = Op->an, am = Op->am;
    const double *x1 = x, *x2 = x + an;
    double *y1 = y, *y2 = y + an;
    linop_mul_t(y1, Op, x2);                               /* :45 */
    PFOR
    for (int64_t i = 0; i < an; i++) y1[i] += x1[i];       /* :46 */
    linop_mul(y2, Op, x1);                                 /* :47 */
    PFOR
    for (int64_t i = 0; i < am; i++) y2[i] -= x2[i];       /* :48 */
}

/* HSDEMatrix mul! (HSDEAffine.jl:131-147): Y = [I Q'; Q -I] B.  Same shape as kkt_mul. */
static void hsdem_mul(double *y, const linop_t *Q, const double *x) { kkt_mul(y, Q, x); }

/* ------------------------------------------------------------------------------------
 * conjugategradient!  (utilities/conjugategradients.jl:31-55)
 * matrix-free on a callback so the same routine serves KKTMatrix, HSDEMatrix and the
 * plain SPD matrix of test/conjugateGradient.jl
 * ---------------------------------------------------------------------------------- */
typedef void (*mulfn_t)(double *y, const void *ctx, const double *x);

static void mul_kkt_cb(double *y, const void *ctx, const double *x) { kkt_mul(y, (const linop_t *)ctx, x); }
static void mul_csc_cb(double *y, const void *ctx, const double *x) { csc_mul(y, (const csc_t *)ctx, x); }

static int64_t conjgrad(double *x, mulfn_t mul, const void *ctx, const double *b, double *r, double *p,
                        double *Ap, int64_t N, double tol, int64_t max_iters)
{
    mul(Ap, ctx, x);                                         /* :32 */
    PFOR
    for (int64_t i = 0; i < N; i++) r[i] = b[i] - Ap[i];     /* :33 */
    PFOR
    for (int64_t i = 0; i < N; i++) p[i] = r[i];             /* :34 */
    double rn = vdot(r, r, N);                               /* :35 */
    int64_t iter = 1;                                        /* :36 */
    for (;;) {
        mul(Ap, ctx, p);                                     /* :38 */
        double alpha = rn / vdot(Ap, p, N);                  /* :39 */
        PFOR
        for (int64_t i = 0; i < N; i++) x[i] += alpha * p[i];   /* :40 */
        PFOR
        for (int64_t i = 0; i < N; i++) r[i] -= alpha * Ap[i];  /* :41 */
        if (vnorm(r, N) <= tol || iter >= max_iters) break;  /* :42 */
        double rnold = rn;                                   /* :45 */
        rn = vdot(r, r, N);                                  /* :46 */
        double beta = rn / rnold;                            /* :47 */
        PFOR
        for (int64_t i = 0; i < N; i++) p[i] *= beta;        /* :49 */
        PFOR
        for (int64_t i = 0; i < N; i++) p[i] += r[i];        /* :50 */
        iter += 1;                                           /* :51 */
    }
    return iter;                                             /* :54 (the @warn at :53 is the caller's flag) */
}

/* ------------------------------------------------------------------------------------
 * AffinePlusLinear  (utilities/affinepluslinear.jl:58-126)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    linop_t op;
    int64_t beta;
    double *b, *q;          /* am, an */
    double *rhs;            /* an+am, second half pre-filled with b (:76-77) */
    int decreasing_accuracy;
    int64_t i;              /* call count, starts at 1 (:78) */
    int64_t cgiter;
    double *r, *p, *z, *xinit; /* CGdata (conjugategradients.jl:1-11) */
    int firstrun;
    int cg_maxit_warned;    /* sticky flag standing in for the @warn at :120 */
} apl_t;

static apl_t *apl_new(linop_t op, const double *b, const double *q, int64_t beta, int decreasing)
{
    apl_t *S = (apl_t *)calloc(1, sizeof(apl_t));
    int64_t am = op.am, an = op.an, N = am + an;
    S->op = op;
    S->beta = beta;
    S->b = (double *)calloc((size_t)am, sizeof(double));
    S->q = (double *)calloc((size_t)an, sizeof(double));
    if (b) memcpy(S->b, b, sizeof(double) * (size_t)am);
    if (q) memcpy(S->q, q, sizeof(double) * (size_t)an);
    S->rhs = (double *)calloc((size_t)N, sizeof(double));
    memcpy(S->rhs + an, S->b, sizeof(double) * (size_t)am);
    S->decreasing_accuracy = decreasing;
    S->i = 1;
    S->cgiter = 0;
    S->r = (double *)calloc((size_t)N, sizeof(double));
    S->p = (double *)calloc((size_t)N, sizeof(double));
    S->z = (double *)calloc((size_t)N, sizeof(double));
    S->xinit = (double *)calloc((size_t)N, sizeof(double));
    S->firstrun = 1;
    return S;
}
static void apl_free(apl_t *S)
{
    if (!S) return;
    free(S->b); free(S->q); free(S->rhs); free(S->r); free(S->p); free(S->z); free(S->xinit);
    free(S);
}

#define JULIA_EPS 2.220446049250313e-16 /* eps(Float64) */

/* prox!(y, S::AffinePlusLinear, x)  (affinepluslinear.jl:83-126) */
static void apl_prox(double *y, apl_t *S, const double *x)
{
    int64_t an = S->op.an, am = S->op.am, N = an + am;
    double *rhs1 = S->rhs;
    const double *x1 = x, *x2 = x + an;
    double *y2 = y + an;
    double beta = (double)S->beta;
    linop_mul_t(rhs1, &S->op, x2);                                                   /* :94 */
    for (int64_t i = 0; i < an; i++) rhs1[i] = beta * rhs1[i] + x1[i] - S->q[i];     /* :95 */
    if (S->firstrun) {                                                               /* :101-104 */
        memcpy(S->xinit, x, sizeof(double) * (size_t)N);
        S->firstrun = 0;
    }
    memcpy(y, S->xinit, sizeof(double) * (size_t)N);                                 /* :106 */
    double tol;
    if (S->decreasing_accuracy) {                                                    /* :108-112 */
        double t = pow(0.2, sqrt((double)S->i));
        double floor_ = (double)an * JULIA_EPS;
        tol = t > floor_ ? t : floor_;
    } else {
        tol = (double)an * JULIA_EPS;
    }
    S->i += 1;                                                                       /* :114 */
    int64_t max_iters = 1000;                                                        /* :115 */
    int64_t iter = conjgrad(y, mul_kkt_cb, &S->op, S->rhs, S->r, S->p, S->z, N, tol, max_iters); /* :117 */
    if (iter == max_iters) S->cg_maxit_warned = 1;                                   /* :120 */
    S->cgiter = iter;                                                                /* :121 */
    memcpy(S->xinit, y, sizeof(double) * (size_t)N);                                 /* :122 */
    for (int64_t i = 0; i < am; i++) y2[i] *= beta;                                  /* :124 */
}

/* HSDEMatrix.prox! (HSDEAffine.jl:105-126): fixed tolerance 2l*eps, then v <- Q u. Own CG state. */
typedef struct {
    linop_t Q;
    double *r, *p, *z, *xinit;
    int firstrun;
} hsdem_t;

static void hsdem_prox(double *y, hsdem_t *M, const double *x)
{
    int64_t l = M->Q.an, N = 2 * l;
    double tol = (double)N * JULIA_EPS;                         /* :106 */
    if (M->firstrun) {                                          /* :109-112 */
        memcpy(M->xinit, x, sizeof(double) * (size_t)N);
        M->firstrun = 0;
    }
    memcpy(y, M->xinit, sizeof(double) * (size_t)N);            /* :114 */
    conjgrad(y, mul_kkt_cb, &M->Q, x, M->r, M->p, M->z, N, tol, 1000); /* :116 */
    memcpy(M->xinit, y, sizeof(double) * (size_t)N);            /* :119 */
    hsdeq_mul(y + l, &M->Q, y);                                 /* :122-124  v = Q*u */
}

/* ------------------------------------------------------------------------------------
 * Cone projections.  The arithmetic is ProximalOperators.jl's (un-vendored); restated
 * from its published algorithms, mapped at cones.jl:4-14.
 * ---------------------------------------------------------------------------------- */

/* cyclic two-sided Jacobi eigendecomposition of a dense symmetric d x d matrix (row-major).
 * Stands in for LAPACK dspev used by IndPSD; the projection V max(L,0) V' is unique, so the
 * choice of eigen-solver does not affect parity (hard parts). */
static void jacobi_eig(double *S, double *V, int64_t d)
{
    for (int64_t i = 0; i < d; i++)
        for (int64_t j = 0; j < d; j++) V[i * d + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int64_t i = 0; i < d; i++) {
            diag += S[i * d + i] * S[i * d + i];
            for (int64_t j = i + 1; j < d; j++) off += S[i * d + j] * S[i * d + j];
        }
        if (off == 0.0 || off <= 1e-34 * (diag + off)) break;
        for (int64_t p = 0; p < d - 1; p++) {
            for (int64_t q = p + 1; q < d; q++) {
                double apq = S[p * d + q];
                if (apq == 0.0) continue;
                double app = S[p * d + p], aqq = S[q * d + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int64_t k = 0; k < d; k++) { /* columns p,q */
                    double skp = S[k * d + p], skq = S[k * d + q];
                    S[k * d + p] = c * skp - s * skq;
                    S[k * d + q] = s * skp + c * skq;
                }
                for (int64_t k = 0; k < d; k++) { /* rows p,q */
                    double spk = S[p * d + k], sqk = S[q * d + k];
                    S[p * d + k] = c * spk - s * sqk;
                    S[q * d + k] = s * spk + c * sqk;
                }
                for (int64_t k = 0; k < d; k++) {
                    double vkp = V[k * d + p], vkq = V[k * d + q];
                    V[k * d + p] = c * vkp - s * vkq;
                    V[k * d + q] = s * vkp + c * vkq;
                }
            }
        }
    }
}

static int64_t sdp_dim(int64_t len)
{
    /* len = d(d+1)/2 */
    int64_t d = (int64_t)floor(sqrt(0.25 + 2.0 * (double)len) - 0.5 + 1e-9);
    return d;
}

/* IndPSD(scaling=true) on the packed lower triangle, column-major (MathProgBase svec):
 * diagonal *= sqrt2, eig, clamp at 0, reassemble, repack, diagonal /= sqrt2. */
static void prox_sdp(double *y, const double *x, int64_t len)
{
    int64_t d = sdp_dim(len);
    double *S = (double *)malloc(sizeof(double) * (size_t)(d * d));
    double *V = (double *)malloc(sizeof(double) * (size_t)(d * d));
    double *P = (double *)malloc(sizeof(double) * (size_t)(d * d));
    const double sq2 = sqrt(2.0);
    int64_t k = 0;
    for (int64_t j = 0; j < d; j++)
        for (int64_t i = j; i < d; i++, k++) {
            double v = x[k];
            if (i == j) v *= sq2;
            S[i * d + j] = v;
            S[j * d + i] = v;
        }
    jacobi_eig(S, V, d);
    for (int64_t i = 0; i < d; i++)
        for (int64_t j = 0; j <= i; j++) {
            double acc = 0.0;
            for (int64_t e = 0; e < d; e++) {
                double lam = S[e * d + e];
                if (lam > 0.0) acc += V[i * d + e] * lam * V[j * d + e];
            }
            P[i * d + j] = acc;
            P[j * d + i] = acc;
        }
    k = 0;
    for (int64_t j = 0; j < d; j++)
        for (int64_t i = j; i < d; i++, k++) {
            double v = P[i * d + j];
            if (i == j) v /= sq2;
            y[k] = v;
        }
    free(S); free(V); free(P);
}

/* IndSOC on (t, w), first entry the scalar */
static void prox_soc(double *y, const double *x, int64_t len)
{
    double nx = vnorm(x + 1, len - 1);
    double t = x[0];
    if (t <= -nx) {
        for (int64_t i = 0; i < len; i++) y[i] = 0.0;
    } else if (t >= nx) {
        for (int64_t i = 0; i < len; i++) y[i] = x[i];
    } else {
        double r = 0.5 * (1.0 + t / nx);
        y[0] = r * nx;
        for (int64_t i = 1; i < len; i++) y[i] = r * x[i];
    }
}

/* IndRotatedSOC of ProximalOperators.jl (published algorithm, restated): rotate (x1, x2) by pi/4 with the
 * literal 0.7071067811865475, project (x1', (x2', w)) onto the SOC, rotate back.  PARITY UNPINNED. */
static void prox_socrot(double *y, const double *x, int64_t len)
{
    const double c = 0.7071067811865475;
    double x1 = c * x[0] + c * x[1];
    double x2 = c * x[0] - c * x[1];
    double nw = vnorm(x + 2, len - 2);
    double nx = sqrt(x2 * x2 + nw * nw);
    double t = x1, y1, y2;
    if (t <= -nx) {
        for (int64_t i = 0; i < len; i++) y[i] = 0.0;
        y1 = 0.0; y2 = 0.0;
    } else if (t >= nx) {
        y1 = x1; y2 = x2;
        for (int64_t i = 2; i < len; i++) y[i] = x[i];
    } else {
        double r = 0.5 * (1.0 + t / nx);
        y1 = r * nx; y2 = r * x2;
        for (int64_t i = 2; i < len; i++) y[i] = r * x[i];
    }
    y[0] = c * y1 + c * y2;
    y[1] = c * y1 - c * y2;
}

/* Exponential cone cl{(r,s,t): s > 0, s exp(r/s) <= t}.  ProximalOperators' IndExpPrimal follows the projection
 * of SCS (cones.c: proj_exp_cone, exp_newton_one_d, exp_calc_grad, exp_get_rho_ub), restated here with its
 * tolerance 1e-8 and iteration caps of 100.  PARITY UNPINNED (un-vendored dependency, no reference test). */
#define EXP_TOL 1e-8
static double exp_newton_one_d(double rho, double y_hat, double z_hat)
{
    double t = fmax(-z_hat, 1e-6);
    for (int i = 0; i < 100; i++) {
        double f = t * (t + z_hat) / rho / rho - y_hat / rho + log(t / rho) + 1.0;
        double fp = (2.0 * t + z_hat) / rho / rho + 1.0 / t;
        t = t - f / fp;
        if (t <= -z_hat) return 0.0;
        else if (t <= 0.0) return z_hat;
        else if (fabs(f) < EXP_TOL) break;
    }
    return t + z_hat;
}
static double exp_calc_grad(const double *v, double *x, double rho)
{
    x[2] = exp_newton_one_d(rho, v[1], v[2]);
    x[1] = (x[2] - v[2]) * x[2] / rho;
    x[0] = v[0] - rho;
    if (x[1] <= 1e-12) return x[0];
    return x[0] + x[1] * log(x[1] / x[2]);
}
static void prox_exp3(double *y, const double *v)
{
    double r = v[0], s = v[1], t = v[2];
    if ((s * exp(r / s) - t <= EXP_TOL && s > 0.0) || (r <= 0.0 && s == 0.0 && t >= 0.0)) {
        y[0] = r; y[1] = s; y[2] = t;
        return;
    }
    if ((-r < 0.0 && r * exp(s / r) + exp(1.0) * t <= EXP_TOL) || (-r == 0.0 && -s >= 0.0 && -t >= 0.0)) {
        y[0] = y[1] = y[2] = 0.0;
        return;
    }
    if (r < 0.0 && s < 0.0) {
        y[0] = r; y[1] = 0.0; y[2] = fmax(t, 0.0);
        return;
    }
    double x[3], lb = 0.0, ub = 0.125;
    while (exp_calc_grad(v, x, ub) > 0.0) { lb = ub; ub *= 2.0; }
    for (int i = 0; i < 100; i++) {
        double rho = (ub + lb) / 2.0;
        double g = exp_calc_grad(v, x, rho);
        if (g > 0.0) lb = rho; else ub = rho;
        if (ub - lb < EXP_TOL) break;
    }
    y[0] = x[0]; y[1] = x[1]; y[2] = x[2];
}
static int prox_exp(double *y, const double *x, int64_t len, int dualcone)
{
    if (len % 3 != 0) return -1;
    for (int64_t k = 0; k < len; k += 3) {
        if (!dualcone) prox_exp3(y + k, x + k);
        else { /* IndExpDual by Moreau: x + P_K(-x) */
            double nv[3] = {-x[k], -x[k + 1], -x[k + 2]}, p[3];
            prox_exp3(p, nv);
            for (int j = 0; j < 3; j++) y[k + j] = x[k + j] + p[j];
        }
    }
    return 0;
}

static int prox_cone(double *y, int type, const double *x, int64_t len)
{
    switch (type) {
    case CONE_SOCROT: if (len < 2) return -1; prox_socrot(y, x, len); return 0;
    case CONE_EXPPRIMAL: return prox_exp(y, x, len, 0);
    case CONE_EXPDUAL: return prox_exp(y, x, len, 1);
    case CONE_FREE: for (int64_t i = 0; i < len; i++) y[i] = x[i]; return 0;
    case CONE_ZERO: for (int64_t i = 0; i < len; i++) y[i] = 0.0; return 0;
    /* NaN passes through (IndNonnegative: `x < 0 ? 0 : x`), as it does through Julia's max(x, 0) below */
    case CONE_NONNEG: for (int64_t i = 0; i < len; i++) y[i] = x[i] < 0.0 ? 0.0 : x[i]; return 0;
    case CONE_NONPOS: for (int64_t i = 0; i < len; i++) y[i] = x[i] > 0.0 ? 0.0 : x[i]; return 0;
    case CONE_SOC: prox_soc(y, x, len); return 0;
    case CONE_SDP: prox_sdp(y, x, len); return 0;
    default: return -1;
    }
}

/* proxDual! (cones.jl:80-85 generic Moreau, :97-102 shortcuts) */
static int prox_cone_dual(double *y, int type, const double *x, int64_t len)
{
    switch (type) {
    case CONE_ZERO: return prox_cone(y, CONE_FREE, x, len);   /* :98 */
    case CONE_FREE: return prox_cone(y, CONE_ZERO, x, len);   /* :100 */
    case CONE_NONNEG: return prox_cone(y, CONE_NONNEG, x, len); /* :101 */
    case CONE_NONPOS: return prox_cone(y, CONE_NONPOS, x, len); /* :102 */
    default: {
        double *neg = (double *)calloc((size_t)(len > 0 ? len : 1), sizeof(double));
        for (int64_t i = 0; i < len; i++) neg[i] = -x[i];
        int rc = prox_cone(y, type, neg, len);                /* :81 */
        for (int64_t i = 0; i < len; i++) y[i] = x[i] + y[i]; /* :82-84 */
        free(neg);
        return rc;
    }
    }
}

/* ConeProduct (cones.jl:31-77): contiguous ranges covering 1:N */
typedef struct {
    int64_t ncones;
    int32_t *types;
    int64_t *lens;
    int64_t total;
} coneprod_t;

static coneprod_t coneprod_new(int64_t nc, const int32_t *types, const int64_t *lens)
{
    coneprod_t K;
    K.ncones = nc;
    K.types = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nc > 0 ? nc : 1));
    K.lens = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nc > 0 ? nc : 1));
    K.total = 0;
    for (int64_t i = 0; i < nc; i++) {
        K.types[i] = types[i];
        K.lens[i] = lens[i];
        K.total += lens[i];
    }
    return K;
}
static void coneprod_free(coneprod_t *K) { free(K->types); free(K->lens); }

/* prox!(y, C::ConeProduct, x) (cones.jl:89-94) / proxDual! (:106-111) */
static int coneprod_prox(double *y, const coneprod_t *K, const double *x, int dual)
{
    int64_t off = 0;
    int rc = 0;
    for (int64_t i = 0; i < K->ncones; i++) {
        rc |= dual ? prox_cone_dual(y + off, K->types[i], x + off, K->lens[i])
                   : prox_cone(y + off, K->types[i], x + off, K->lens[i]);
        off += K->lens[i];
    }
    return rc;
}

/* DualConeProduct prox! (cones.jl:122-142): K2 x K1* x R+ x K2* x K1 x R+ */
static int dualconeprod_prox(double *y, const coneprod_t *K1, const coneprod_t *K2, const double *x)
{
    int64_t m = K1->total, n = K2->total, nu = n + m + 1;
    int rc = 0;
    rc |= coneprod_prox(y, K2, x, 0);                          /* :136 */
    rc |= coneprod_prox(y + n, K1, x + n, 1);                  /* :137 */
    y[nu - 1] = x[nu - 1] < 0.0 ? 0.0 : x[nu - 1];             /* :138  max(tau, 0), NaN stays NaN */
    rc |= coneprod_prox(y + nu, K2, x + nu, 1);                /* :139 */
    rc |= coneprod_prox(y + nu + n, K1, x + nu + n, 0);        /* :140 */
    y[2 * nu - 1] = x[2 * nu - 1] < 0.0 ? 0.0 : x[2 * nu - 1]; /* :141 */
    return rc;
}

/* ------------------------------------------------------------------------------------
 * Model: HSDE conic form (problemforms/HSDE/HSDE.jl:7-29) or Feasibility form
 * (problemforms/Feasibility/Feasibility.jl) with S1 = AffinePlusLinear, S2 = ConeProduct
 * ---------------------------------------------------------------------------------- */
#define REC_LEN 10 /* i, p, d, g, ctx, bty, kappa, tau, cgiter, status   (HSDE)
                      i, err, 0...,              cgiter, status           (Feasibility) */

typedef struct {
    int form; /* 0 = HSDE conic, 1 = Feasibility */
    csc_t *A;
    double *b, *c; /* HSDE data */
    coneprod_t K1, K2; /* HSDE: K1 over m rows, K2 over n vars. Feasibility: K1 = S2 over N */
    apl_t *S1;
    int64_t N; /* iterate length */
    /* algorithm */
    int alg;
    double alpha, alpha1, alpha2, betaA;
    int64_t iproj;
    /* algorithm data (GAPData / GAPAData / FISTAData / DykstraData / GAPPData) */
    double *x, *tmp1, *tmp2;
    double alpha12;           /* gapa.jl:29 */
    double fista_t;           /* fista.jl:24 */
    double *fy, *fxold;       /* fista.jl:13-20 */
    double *dp, *dq, *dy;     /* dykstra.jl:12-18 */
    double *work1, *work2, *work3; /* temporaries (allocated per step in the reference) */
    /* status (HSDEStatus.jl:2-16 / FeasibilityStatus.jl) */
    int status;
    int checked;
    int64_t checki;
    double eps;
    int64_t cur_i;
    double *prev; /* FeasibilityStatus.prev, NaN-initialised (Feasibility.jl:79) */
    /* history sink */
    double *hist;
    int64_t hist_cap, hist_len;
    /* GAPP line-search log: alpha_best of each projected iteration */
    double last_alphabest;
    /* direct = true (HSDE.jl:10-15): S1 = IndAffine([Q -I], 0), exact projection through a dense
       Cholesky factor of I + Q Q' (ProximalOperators factorises the same normal equations) */
    int direct;
    double *chol; /* l x l, lower triangle, row-major */
    /* IndBox(lo, hi) entries of S2 in the Feasibility form (test/testfeasibility.jl:10) */
    double *box_lo, *box_hi;
    unsigned char *box_on;
    /* LineSearchWrapper (wrappers/linesearch.jl): 0 = no wrapper */
    int64_t lsinterval;
    double *ls1, *ls2, *ls3, *lsres; /* LineSearchWrapperData tmp1, tmp2, tmp3, res (:9-17) */
} model_t;

static void model_alloc_vectors(model_t *M)
{
    size_t N = (size_t)M->N;
    M->x = (double *)calloc(N, sizeof(double));
    M->tmp1 = (double *)calloc(N, sizeof(double));
    M->tmp2 = (double *)calloc(N, sizeof(double));
    M->fy = (double *)calloc(N, sizeof(double));
    M->fxold = (double *)calloc(N, sizeof(double));
    M->dp = (double *)calloc(N, sizeof(double));
    M->dq = (double *)calloc(N, sizeof(double));
    M->dy = (double *)calloc(N, sizeof(double));
    M->work1 = (double *)calloc(N, sizeof(double));
    M->work2 = (double *)calloc(N, sizeof(double));
    M->work3 = (double *)calloc(N, sizeof(double));
    M->prev = (double *)calloc(N, sizeof(double));
    for (size_t i = 0; i < N; i++) M->prev[i] = NAN;
    M->alpha12 = 2.0;
    M->fista_t = 1.0;
    M->alg = ALG_GAP;
    M->alpha = 0.8; M->alpha1 = 1.8; M->alpha2 = 1.8; M->betaA = 0.0; M->iproj = 100;
    M->status = ST_CONTINUE;
    M->checki = 100;
    M->eps = 1e-5;
}

FOSOR_API void *fosor_create_conic(int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval,
                                   const double *nzval, int64_t index_base, const double *b, const double *c,
                                   int64_t nc1, const int32_t *types1, const int64_t *lens1, int64_t nc2,
                                   const int32_t *types2, const int64_t *lens2)
{
    model_t *M = (model_t *)calloc(1, sizeof(model_t));
    M->form = 0;
    M->A = csc_new(m, n, colptr, rowval, nzval, index_base);
    M->b = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    M->c = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    memcpy(M->b, b, sizeof(double) * (size_t)m);
    memcpy(M->c, c, sizeof(double) * (size_t)n);
    M->K1 = coneprod_new(nc1, types1, lens1);
    M->K2 = coneprod_new(nc2, types2, lens2);
    if (M->K1.total != m || M->K2.total != n) { /* cones.jl:66-72 coverage asserts */
        csc_free(M->A); free(M->b); free(M->c); coneprod_free(&M->K1); coneprod_free(&M->K2); free(M);
        return NULL;
    }
    /* HSDE.jl:17-22: S1 = AffinePlusLinear(Q, 0, 0, 1, decreasing_accuracy=true) */
    linop_t Q;
    Q.kind = 1; Q.A = M->A; Q.b = M->b; Q.c = M->c; Q.am = m + n + 1; Q.an = m + n + 1;
    M->S1 = apl_new(Q, NULL, NULL, 1, 1);
    M->N = 2 * (m + n + 1); /* HSDE.jl:28 */
    model_alloc_vectors(M);
    /* HSDE_getinitialvalue (HSDE.jl:40-47) */
    int64_t l = m + n + 1;
    M->x[l - 1] = 1.0;
    M->x[2 * l - 1] = 1.0;
    return M;
}

/* Feasibility(S1 = AffinePlusLinear(A,b,q,beta; decreasing_accuracy), S2 = ConeProduct, N = an+am) */
/* Same model on CSC arrays that stay with the caller (0-based, must outlive the model): a dense 20000 x 40000
 * matrix is 12.8 GB as SparseMatrixCSC{Float64,Int64} (types.jl:35) and is not copied a second time. */
FOSOR_API void *fosor_create_conic_borrowed(int64_t m, int64_t n, int64_t *colptr, int64_t *rowval, double *nzval,
                                            const double *b, const double *c, int64_t nc1, const int32_t *types1,
                                            const int64_t *lens1, int64_t nc2, const int32_t *types2,
                                            const int64_t *lens2)
{
    model_t *M = (model_t *)calloc(1, sizeof(model_t));
    M->form = 0;
    M->A = csc_borrow(m, n, colptr, rowval, nzval);
    M->b = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    M->c = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    memcpy(M->b, b, sizeof(double) * (size_t)m);
    memcpy(M->c, c, sizeof(double) * (size_t)n);
    M->K1 = coneprod_new(nc1, types1, lens1);
    M->K2 = coneprod_new(nc2, types2, lens2);
    if (M->K1.total != m || M->K2.total != n) {
        csc_free(M->A); free(M->b); free(M->c); coneprod_free(&M->K1); coneprod_free(&M->K2); free(M);
        return NULL;
    }
    linop_t Q;
    Q.kind = 1; Q.A = M->A; Q.b = M->b; Q.c = M->c; Q.am = m + n + 1; Q.an = m + n + 1;
    M->S1 = apl_new(Q, NULL, NULL, 1, 1);
    M->N = 2 * (m + n + 1);
    model_alloc_vectors(M);
    int64_t l = m + n + 1;
    M->x[l - 1] = 1.0;
    M->x[2 * l - 1] = 1.0;
    return M;
}

/* Synthetic dense matrix of bench.py's config 2 written straight into CSC arrays (every entry stored, row
 * indices 0..m-1 per column): A[i,j] = scale * N(0,1), counter-based generator (splitmix64 + Box-Muller) so that
 * the fill can run on all host threads.  Not part of the reference: input synthesis for the CPU baseline. */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
FOSOR_API void fosor_gen_dense_csc(int64_t m, int64_t n, uint64_t seed, double scale, int64_t *colptr,
                                   int64_t *rowval, double *nzval)
{
    for (int64_t j = 0; j <= n; j++) colptr[j] = j * m;
    PFOR
    for (int64_t j = 0; j < n; j++) {
        for (int64_t i = 0; i < m; i++) {
            const uint64_t ctr = (uint64_t)(j * m + i);
            const uint64_t u1 = splitmix64(seed * 0x100000001B3ULL + 2 * ctr);
            const uint64_t u2 = splitmix64(seed * 0x100000001B3ULL + 2 * ctr + 1);
            const double a = ((double)(u1 >> 11) + 1.0) * (1.0 / 9007199254740993.0); /* (0,1) */
            const double b = (double)(u2 >> 11) * (1.0 / 9007199254740992.0);
            nzval[j * m + i] = scale * sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);
            rowval[j * m + i] = i;
        }
    }
}
/* y = A x (transpose = 0) or A' x (transpose = 1) on caller-owned 0-based CSC arrays */
FOSOR_API void fosor_csc_mul(int64_t m, int64_t n, int64_t *colptr, int64_t *rowval, double *nzval, const double *x,
                             double *y, int32_t transpose)
{
    csc_t A;
    A.m = m; A.n = n; A.colptr = colptr; A.rowval = rowval; A.nzval = nzval; A.borrowed = 1;
    if (transpose) csc_mul_t(y, &A, x);
    else csc_mul(y, &A, x);
}

FOSOR_API void *fosor_create_feasibility(int64_t am, int64_t an, const int64_t *colptr, const int64_t *rowval,
                                         const double *nzval, int64_t index_base, const double *b,
                                         const double *q, int64_t beta, int32_t decreasing, int64_t nc,
                                         const int32_t *types, const int64_t *lens)
{
    model_t *M = (model_t *)calloc(1, sizeof(model_t));
    M->form = 1;
    M->A = csc_new(am, an, colptr, rowval, nzval, index_base);
    M->K1 = coneprod_new(nc, types, lens);
    M->K2 = coneprod_new(0, NULL, NULL);
    if (M->K1.total != am + an) {
        csc_free(M->A); coneprod_free(&M->K1); coneprod_free(&M->K2); free(M);
        return NULL;
    }
    linop_t Op;
    Op.kind = 0; Op.A = M->A; Op.b = NULL; Op.c = NULL; Op.am = am; Op.an = an;
    M->S1 = apl_new(Op, b, q, beta, decreasing);
    M->N = am + an;
    model_alloc_vectors(M); /* getinitialvalue = zeros (Feasibility.jl:57-58) */
    return M;
}

FOSOR_API void fosor_destroy(void *h)
{
    model_t *M = (model_t *)h;
    if (!M) return;
    csc_free(M->A);
    free(M->b); free(M->c);
    coneprod_free(&M->K1); coneprod_free(&M->K2);
    apl_free(M->S1);
    free(M->x); free(M->tmp1); free(M->tmp2); free(M->fy); free(M->fxold);
    free(M->dp); free(M->dq); free(M->dy); free(M->work1); free(M->work2); free(M->work3); free(M->prev);
    free(M->chol);
    free(M->ls1); free(M->ls2); free(M->ls3); free(M->lsres);
    free(M->box_on); free(M->box_lo); free(M->box_hi);
    free(M);
}

FOSOR_API int64_t fosor_iterate_length(void *h) { return ((model_t *)h)->N; }

/* algorithm constructors (a20): GAP gap.jl:6-13, GAPA gapa.jl:9-15, FISTA fista.jl:6-11,
 * Dykstra dykstra.jl:6-10, GAPP gapproj.jl:6-14.  DR/AP are GAP(alpha,2,2)/GAP(alpha,1,1). */
FOSOR_API void fosor_set_algorithm(void *h, int32_t alg, double alpha, double alpha1, double alpha2, double beta,
                                   int64_t iproj)
{
    model_t *M = (model_t *)h;
    M->alg = alg;
    M->alpha = alpha; M->alpha1 = alpha1; M->alpha2 = alpha2; M->betaA = beta; M->iproj = iproj;
    M->alpha12 = 2.0; /* gapa.jl:29 */
    M->fista_t = 1.0; /* fista.jl:24 */
    memset(M->fy, 0, sizeof(double) * (size_t)M->N);
    memset(M->fxold, 0, sizeof(double) * (size_t)M->N);
    memset(M->dp, 0, sizeof(double) * (size_t)M->N); /* dykstra.jl:22 */
    memset(M->dq, 0, sizeof(double) * (size_t)M->N);
}

/* LineSearchWrapper(alg; lsinterval) (wrappers/linesearch.jl:19-24): GAP and GAPA only (support_linesearch) */
FOSOR_API int32_t fosor_set_linesearch(void *h, int64_t lsinterval)
{
    model_t *M = (model_t *)h;
    if (lsinterval < 0) return -1;
    M->lsinterval = lsinterval;
    if (lsinterval > 0 && !M->ls1) {
        size_t N = (size_t)M->N;
        M->ls1 = (double *)calloc(N, sizeof(double));
        M->ls2 = (double *)calloc(N, sizeof(double));
        M->ls3 = (double *)calloc(N, sizeof(double));
        M->lsres = (double *)calloc(N, sizeof(double));
    }
    return 0;
}

FOSOR_API void fosor_set_iterate(void *h, const double *z)
{
    model_t *M = (model_t *)h;
    memcpy(M->x, z, sizeof(double) * (size_t)M->N);
}
FOSOR_API void fosor_get_iterate(void *h, double *z)
{
    model_t *M = (model_t *)h;
    memcpy(z, M->x, sizeof(double) * (size_t)M->N);
}
/* which: 0 x, 1 tmp1, 2 tmp2, 3 S1.xinit, 4 S1.rhs, 6 FISTA y, 7 Dykstra p, 8 Dykstra q */
static double *state_ptr(model_t *M, int32_t which)
{
    switch (which) {
    case 1: return M->tmp1;
    case 2: return M->tmp2;
    case 3: return M->S1->xinit;
    case 4: return M->S1->rhs;
    case 6: return M->fy;
    case 7: return M->dp;
    case 8: return M->dq;
    default: return M->x;
    }
}
FOSOR_API void fosor_get_state(void *h, int32_t which, double *out)
{
    model_t *M = (model_t *)h;
    memcpy(out, state_ptr(M, which), sizeof(double) * (size_t)M->N);
}
/* Restores a persistent vector / scalar (lock-step tests put several restatements into one common state):
 * which as in fosor_get_state; setting 3 (xinit) also clears S1's first-run flag (affinepluslinear.jl:101-104). */
FOSOR_API void fosor_set_state(void *h, int32_t which, const double *in)
{
    model_t *M = (model_t *)h;
    memcpy(state_ptr(M, which), in, sizeof(double) * (size_t)M->N);
    if (which == 3) M->S1->firstrun = 0;
}
/* which: 0 S1.i, 2 alpha12, 3 FISTA t */
FOSOR_API void fosor_set_scalar(void *h, int32_t which, double v)
{
    model_t *M = (model_t *)h;
    if (which == 0) M->S1->i = (int64_t)v;
    else if (which == 2) M->alpha12 = v;
    else if (which == 3) M->fista_t = v;
}
/* 0 = restatement, 1 = long-double reductions (FOS_ORACLE_HP), 2 = threaded (FOS_ORACLE_MT) */
FOSOR_API int32_t fosor_variant(void)
{
#if defined(FOS_ORACLE_HP)
    return 1;
#elif defined(FOS_ORACLE_MT)
    return 2;
#else
    return 0;
#endif
}
FOSOR_API int32_t fosor_threads(void)
{
#ifdef FOS_ORACLE_MT
    return (int32_t)omp_get_max_threads();
#else
    return 1;
#endif
}
FOSOR_API double fosor_get_fista_t(void *h) { return ((model_t *)h)->fista_t; }
FOSOR_API int64_t fosor_get_s1_calls(void *h) { return ((model_t *)h)->S1->i; }
FOSOR_API int64_t fosor_get_cgiter(void *h) { return ((model_t *)h)->S1->cgiter; }
FOSOR_API double fosor_get_alpha12(void *h) { return ((model_t *)h)->alpha12; }
FOSOR_API double fosor_get_alphabest(void *h) { return ((model_t *)h)->last_alphabest; }
FOSOR_API int32_t fosor_get_cg_warned(void *h) { return ((model_t *)h)->S1->cg_maxit_warned; }

/* prox on S2 */
static void prox_S2(model_t *M, double *y, const double *x)
{
    if (M->form == 0) dualconeprod_prox(y, &M->K1, &M->K2, x);
    else {
        coneprod_prox(y, &M->K1, x, 0);
        if (M->box_on)
            for (int64_t i = 0; i < M->N; i++)
                if (M->box_on[i]) y[i] = x[i] < M->box_lo[i] ? M->box_lo[i] : (x[i] > M->box_hi[i] ? M->box_hi[i] : x[i]);
    }
}

/* IndBox(lo, hi) on entries [start, start+len) of the Feasibility iterate */
FOSOR_API int32_t fosor_set_box(void *h, int64_t start, int64_t len, double lo, double hi)
{
    model_t *M = (model_t *)h;
    if (M->form != 1 || start < 0 || len < 0 || start + len > M->N || lo > hi) return -1;
    if (!M->box_on) {
        M->box_on = (unsigned char *)calloc((size_t)M->N, 1);
        M->box_lo = (double *)calloc((size_t)M->N, sizeof(double));
        M->box_hi = (double *)calloc((size_t)M->N, sizeof(double));
    }
    for (int64_t i = start; i < start + len; i++) { M->box_on[i] = 1; M->box_lo[i] = lo; M->box_hi[i] = hi; }
    return 0;
}
/* prox of IndAffine(B, 0) with B = [Q -I] (HSDE.jl:10-15; ProximalOperators' IndAffine:
 * y = z - B'(B B')^{-1} B z, B B' = Q Q' + I):   w = Q u - v ;  (I + Q Q') t = w ;  y = [u - Q't ; v + t] */
static void direct_prox(model_t *M, double *y, const double *x)
{
    const linop_t *Q = &M->S1->op;
    int64_t l = Q->an;
    double *w = (double *)calloc((size_t)l, sizeof(double)), *t = (double *)calloc((size_t)l, sizeof(double));
    linop_mul(w, Q, x);
    for (int64_t i = 0; i < l; i++) w[i] -= x[l + i];
    const double *L = M->chol;
    for (int64_t i = 0; i < l; i++) { /* forward substitution L z = w */
        double s = w[i];
        for (int64_t k = 0; k < i; k++) s -= L[i * l + k] * t[k];
        t[i] = s / L[i * l + i];
    }
    for (int64_t i = l - 1; i >= 0; i--) { /* back substitution L' t = z */
        double s = t[i];
        for (int64_t k = i + 1; k < l; k++) s -= L[k * l + i] * t[k];
        t[i] = s / L[i * l + i];
    }
    linop_mul_t(w, Q, t);
    for (int64_t i = 0; i < l; i++) {
        y[i] = x[i] - w[i];
        y[l + i] = x[l + i] + t[i];
    }
    free(w);
    free(t);
}
static void prox_S1(model_t *M, double *y, const double *x)
{
    if (M->direct) direct_prox(M, y, x);
    else apl_prox(y, M->S1, x);
}

/* Switches the conic model to direct = true / false.  Returns 0, or -1 when the factorisation fails. */
FOSOR_API int32_t fosor_set_direct(void *h, int32_t on)
{
    model_t *M = (model_t *)h;
    if (M->form != 0) return -1;
    free(M->chol);
    M->chol = NULL;
    M->direct = 0;
    if (!on) return 0;
    const linop_t *Q = &M->S1->op;
    int64_t l = Q->an;
    double *Qd = (double *)calloc((size_t)(l * l), sizeof(double)); /* Qd[j*l + i] = Q_ij (column j contiguous) */
    double *e = (double *)calloc((size_t)l, sizeof(double));
    for (int64_t j = 0; j < l; j++) {
        e[j] = 1.0;
        linop_mul(Qd + j * l, Q, e);
        e[j] = 0.0;
    }
    double *G = (double *)calloc((size_t)(l * l), sizeof(double));
    for (int64_t j = 0; j < l; j++)
        for (int64_t i = 0; i < l; i++) {
            double qij = Qd[j * l + i];
            if (qij == 0.0) continue;
            for (int64_t k = 0; k <= i; k++) G[i * l + k] += qij * Qd[j * l + k];
        }
    for (int64_t i = 0; i < l; i++) G[i * l + i] += 1.0;
    for (int64_t j = 0; j < l; j++) { /* Cholesky, lower, in place */
        double d = G[j * l + j];
        for (int64_t k = 0; k < j; k++) d -= G[j * l + k] * G[j * l + k];
        if (!(d > 0.0)) { free(Qd); free(e); free(G); return -1; }
        d = sqrt(d);
        G[j * l + j] = d;
        for (int64_t i = j + 1; i < l; i++) {
            double s2 = G[i * l + j];
            for (int64_t k = 0; k < j; k++) s2 -= G[i * l + k] * G[j * l + k];
            G[i * l + j] = s2 / d;
        }
    }
    free(Qd);
    free(e);
    M->chol = G;
    M->direct = 1;
    return 0;
}

static void push_rec(model_t *M, const double *rec)
{
    if (M->hist && M->hist_len < M->hist_cap) {
        memcpy(M->hist + M->hist_len * REC_LEN, rec, sizeof(double) * REC_LEN);
    }
    M->hist_len += 1;
}

/* checkstatus(::HSDEStatus, z) (HSDEStatus.jl:27-71) */
static int checkstatus_hsde(model_t *M, const double *z, int override)
{
    if (!(M->cur_i % M->checki == 0 || override)) { /* :29, :66-69 */
        M->checked = 0;
        return 0;
    }
    const csc_t *A = M->A;
    int64_t m = A->m, n = A->n, nu = n + m + 1;
    const double *x = z, *y = z + n, *r = z + nu, *s = z + nu + n; /* getvalues :93-102 */
    double tau = z[nu - 1], kappa = z[2 * nu - 1];
    double eps = M->eps;
    double *Ax = M->work1;       /* m */
    double *Aty = M->work2;      /* n */
    double *tmp = M->work3;
    double nb = vnorm(M->b, m), ncn = vnorm(M->c, n);
    /* p = norm(A*x/tau + s/tau - b)/norm(1+norm(b))   :34 */
    csc_mul(Ax, A, x);
    for (int64_t i = 0; i < m; i++) tmp[i] = Ax[i] / tau + s[i] / tau - M->b[i];
    double p = vnorm(tmp, m) / fabs(1.0 + nb);
    /* d = norm(A'*y/tau + c - r/tau)/norm(1+norm(c))   :35 */
    csc_mul_t(Aty, A, y);
    for (int64_t j = 0; j < n; j++) tmp[j] = Aty[j] / tau + M->c[j] - r[j] / tau;
    double d = vnorm(tmp, n) / fabs(1.0 + ncn);
    double ctx = vdot(M->c, x, n); /* :36 */
    double bty = vdot(M->b, y, m); /* :37 */
    double g = fabs(ctx / tau + bty / tau) / (1.0 + fabs(ctx / tau) + fabs(bty / tau)); /* :38 */
    int status = ST_CONTINUE;
    if (p <= eps * (1.0 + nb) && d <= eps * (1.0 + ncn) &&
        g <= eps * (1.0 + fabs(ctx / tau) + fabs(bty / tau))) { /* :54 */
        status = ST_OPTIMAL;
    } else {
        /* :59  norm(A*x + s) <= eps*(-ctx/norm(c)) */
        for (int64_t i = 0; i < m; i++) tmp[i] = Ax[i] + s[i];
        double lhs_u = vnorm(tmp, m);
        if (lhs_u <= eps * (-ctx / ncn)) status = ST_UNBOUNDED;
        else {
            double lhs_i = vnorm(Aty, n); /* :61 */
            if (lhs_i <= eps * (-bty / nb)) status = ST_INFEASIBLE;
        }
    }
    double rec[REC_LEN] = {(double)M->cur_i, p, d, g, ctx, bty, kappa, tau, (double)M->S1->cgiter, (double)status};
    push_rec(M, rec);
    M->status = status; /* :64 */
    M->checked = 1;     /* :65 */
    return 1;
}

/* checkstatus(::FeasibilityStatus, z) (FeasibilityStatus.jl:32-72) */
static int checkstatus_feas(model_t *M, const double *z, int override)
{
    int64_t N = M->N;
    if (M->cur_i % M->checki == 0 || override) {
        acc_t s = 0.0;
        for (int64_t i = 0; i < N; i++) {
            double dlt = M->prev[i] - z[i];
            s += (acc_t)dlt * (acc_t)dlt;
        }
        double err = sqrt((double)s); /* :39 */
        int status = ST_CONTINUE;
        if (err <= M->eps) status = ST_OPTIMAL; /* :55 */
        double rec[REC_LEN] = {(double)M->cur_i, err, 0, 0, 0, 0, 0, 0, (double)M->S1->cgiter, (double)status};
        push_rec(M, rec);
        M->status = status;
        M->checked = 1;
        memcpy(M->prev, z, sizeof(double) * (size_t)N); /* :65 */
        return 1;
    }
    M->checked = 0;
    memcpy(M->prev, z, sizeof(double) * (size_t)N); /* :69 */
    return 0;
}
static int checkstatus(model_t *M, const double *z, int override)
{
    return M->form == 0 ? checkstatus_hsde(M, z, override) : checkstatus_feas(M, z, override);
}

/* GAP.step (gap.jl:42-80) with explicit relaxation parameters so GAPA can share it */
static void step_gap_like(model_t *M, double a1, double a2)
{
    int64_t N = M->N;
    double *x = M->x, *t1 = M->tmp1, *t2 = M->tmp2;
    prox_S1(M, t1, x);                                                         /* gap.jl:45 */
    for (int64_t i = 0; i < N; i++) t1[i] = a1 * t1[i] + (1 - a1) * x[i];      /* :48 */
    prox_S2(M, t2, t1);                                                        /* :55 */
    checkstatus(M, t2, 0);                                                     /* :56 */
    for (int64_t i = 0; i < N; i++) t2[i] = a2 * t2[i] + (1 - a2) * t1[i];     /* :58 */
}

static void step_gap(model_t *M)
{
    int64_t N = M->N;
    double a = M->alpha;
    step_gap_like(M, M->alpha1, M->alpha2);
    for (int64_t i = 0; i < N; i++) M->x[i] = a * M->tmp2[i] + (1 - a) * M->x[i]; /* gap.jl:78 */
}

/* GAPA.step (gapa.jl:80-105) */
static void step_gapa(model_t *M)
{
    int64_t N = M->N;
    double a = M->alpha, bt = M->betaA, a12 = M->alpha12;
    double *x = M->x, *t1 = M->tmp1, *t2 = M->tmp2;
    step_gap_like(M, a12, a12);
    /* normedScalar(tmp2,tmp1,tmp1,x) gapa.jl:36-47 */
    acc_t sum_ = 0.0, n1_ = 0.0, n2_ = 0.0;
    for (int64_t i = 0; i < N; i++) {
        double d1 = t2[i] - t1[i];
        double d2 = t1[i] - x[i];
        sum_ += (acc_t)d1 * (acc_t)d2;
        n1_ += (acc_t)d1 * (acc_t)d1;
        n2_ += (acc_t)d2 * (acc_t)d2;
    }
    double sum = (double)sum_, n1 = (double)n1_, n2 = (double)n2_;
    double scl = fabs(sum) / sqrt(n1 * n2);
    /* clamp(scl, 0, 1) :96 ; isnan -> 0 :97 (clamp propagates NaN) */
    if (scl < 0.0) scl = 0.0;
    if (scl > 1.0) scl = 1.0;
    if (isnan(scl)) scl = 0.0;
    double s = sqrt(1 - scl * scl);        /* :98 */
    double aopt = 2 / (1 + s);             /* :100 */
    M->alpha12 = (1 - bt) * aopt + bt * 2.0; /* :101 */
    for (int64_t i = 0; i < N; i++) x[i] = a * t2[i] + (1 - a) * x[i]; /* :103 */
}

/* FISTA.step (fista.jl:28-48) */
static void step_fista(model_t *M)
{
    int64_t N = M->N;
    double a = M->alpha;
    double *x = M->x, *y = M->fy, *xold = M->fxold, *t1 = M->tmp1;
    if (M->cur_i == 1) memcpy(y, x, sizeof(double) * (size_t)N);            /* :31-33 */
    prox_S1(M, t1, y);                                                      /* :35 */
    for (int64_t i = 0; i < N; i++) t1[i] = a * t1[i] + (1 - a) * y[i];     /* :37 */
    memcpy(xold, x, sizeof(double) * (size_t)N);                            /* :39 */
    prox_S2(M, x, t1);                                                      /* :40 */
    checkstatus(M, x, 0);                                                   /* :41 */
    double told = M->fista_t;                                               /* :44 */
    M->fista_t = (1 + sqrt(1 + 4 * told * told)) / 2;                       /* :45 */
    double coef = (told - 1) / M->fista_t;
    for (int64_t i = 0; i < N; i++) y[i] = x[i] + coef * (x[i] - xold[i]);  /* :46 */
}

/* Dykstra.step (dykstra.jl:26-37) */
static void step_dykstra(model_t *M)
{
    int64_t N = M->N;
    double *x = M->x, *p = M->dp, *q = M->dq, *y = M->dy, *w = M->tmp1;
    for (int64_t i = 0; i < N; i++) w[i] = x[i] + p[i];
    prox_S1(M, y, w);                                            /* :29 */
    for (int64_t i = 0; i < N; i++) p[i] = x[i] + p[i] - y[i];   /* :31 */
    for (int64_t i = 0; i < N; i++) w[i] = y[i] + q[i];
    prox_S2(M, x, w);                                            /* :32 */
    checkstatus(M, x, 0);                                        /* :33 */
    for (int64_t i = 0; i < N; i++) q[i] = y[i] + q[i] - x[i];   /* :35 */
}

/* GAPP.step (gapproj.jl:29-74) */
static void step_gapp(model_t *M)
{
    int64_t N = M->N;
    double a = M->alpha, a1 = M->alpha1, a2 = M->alpha2;
    double *x = M->x, *t1 = M->tmp1, *t2 = M->tmp2;
    prox_S1(M, t1, x);                                                       /* :33 */
    if (M->cur_i % M->iproj == 0) {                                          /* :34 */
        double *t3 = M->work1, *t4 = M->work2, *res = M->fy; /* fy unused by GAPP */
        prox_S2(M, t2, t1);                                                  /* :39 */
        prox_S1(M, res, t2);                                                 /* :40 */
        for (int64_t i = 0; i < N; i++) res[i] = res[i] - t1[i];             /* :41 */
        double normbest = INFINITY, abest = -1.0;
        for (int k = 0; k <= 20; k++) {                                      /* :46 */
            double at = ldexp(1.0, k);                                       /* :47 */
            for (int64_t i = 0; i < N; i++) t3[i] = t1[i] + at * res[i];     /* :48 */
            prox_S2(M, t4, t3);                                              /* :49 */
            double s = 0.0;
            for (int64_t i = 0; i < N; i++) { double dl = t4[i] - t3[i]; s += dl * dl; }
            double nt = sqrt(s);                                             /* :50 */
            if (nt < normbest) { abest = at; normbest = nt; }                /* :52-55 */
        }
        M->last_alphabest = abest;
        for (int64_t i = 0; i < N; i++) t1[i] = t1[i] + abest * res[i];      /* :58 */
        prox_S2(M, t2, t1);                                                  /* :59 */
        checkstatus(M, t2, 0);                                               /* :60 */
        for (int64_t i = 0; i < N; i++) t2[i] = a2 * t2[i] + (1 - a2) * t1[i]; /* :61 */
        memcpy(x, t2, sizeof(double) * (size_t)N);                           /* :62 */
    } else {
        for (int64_t i = 0; i < N; i++) t1[i] = a1 * t1[i] + (1 - a1) * x[i]; /* :64 */
        prox_S2(M, t2, t1);                                                  /* :66 */
        checkstatus(M, t2, 0);                                               /* :67 */
        for (int64_t i = 0; i < N; i++) t2[i] = a2 * t2[i] + (1 - a2) * t1[i]; /* :68 */
        for (int64_t i = 0; i < N; i++) x[i] = a * t2[i] + (1 - a) * x[i];   /* :70 */
    }
}

/* relaxed S1! / S2! of GAP and GAPA (gap.jl:42-60, gapa.jl:61-79); with_status = the real status object
 * (checkstatus inside S2!) or NoStatus (status.jl:3-9) */
static void ls_S1(model_t *M, double *y, const double *x)
{
    double a1 = M->alg == ALG_GAPA ? M->alpha12 : M->alpha1;
    prox_S1(M, y, x);
    for (int64_t i = 0; i < M->N; i++) y[i] = a1 * y[i] + (1 - a1) * x[i];
}
static void ls_S2(model_t *M, double *y, const double *x, int with_status)
{
    double a2 = M->alg == ALG_GAPA ? M->alpha12 : M->alpha2;
    prox_S2(M, y, x);
    if (with_status) checkstatus(M, y, 0);
    for (int64_t i = 0; i < M->N; i++) y[i] = a2 * y[i] + (1 - a2) * x[i];
}

/* step(::LineSearchWrapper, ...) on a line-search iteration (wrappers/linesearch.jl:42-71).  The 33
 * println calls of the reference are not restated. */
static void step_linesearch(model_t *M)
{
    int64_t N = M->N;
    double *x = M->x, *t1 = M->ls1, *t2 = M->ls2, *t3 = M->ls3, *res = M->lsres;
    memcpy(t1, x, sizeof(double) * (size_t)N);                       /* :43 */
    ls_S1(M, t2, x);                                                 /* :47 */
    ls_S2(M, x, t2, 1);                                              /* :48 */
    for (int64_t i = 0; i < N; i++) res[i] = x[i] - t1[i];           /* :51 */
    double best = INFINITY, abest = 1.0, a = 0.1;                    /* :55-57 */
    for (int k = 0; k <= 30; k++) {                                  /* :58 */
        a = a * 1.8;                                                 /* :59 */
        for (int64_t i = 0; i < N; i++) x[i] = t1[i] + a * res[i];   /* :60 */
        ls_S1(M, t2, x);                                             /* :62 */
        ls_S2(M, t3, t2, 0);                                         /* :63 */
        double sacc = 0.0;                                           /* normdiff :77-85 */
        for (int64_t j = 0; j < N; j++) sacc += (x[j] - t3[j]) * (x[j] - t3[j]);
        double testres = sqrt(sacc);
        if (testres < best) {                                        /* :66-69 */
            best = testres;
            abest = a;
        }
    }
    for (int64_t i = 0; i < N; i++) x[i] = t1[i] + abest * res[i];   /* :72 */
    M->last_alphabest = abest;
}

static void do_step(model_t *M)
{
    if (M->lsinterval > 0 && (M->alg == ALG_GAP || M->alg == ALG_GAPA) && M->cur_i % M->lsinterval == 0) {
        step_linesearch(M);                                          /* :41-42 */
        return;
    }
    switch (M->alg) {
    case ALG_GAP: step_gap(M); break;
    case ALG_GAPA: step_gapa(M); break;
    case ALG_FISTA: step_fista(M); break;
    case ALG_DYKSTRA: step_dykstra(M); break;
    case ALG_GAPP: step_gapp(M); break;
    default: break;
    }
}

/* getsol (gap.jl:82-87 and the identical methods of the other algorithms): P2(P1(x)) */
static void do_getsol(model_t *M, double *guess)
{
    prox_S1(M, M->tmp1, M->x);
    prox_S2(M, guess, M->tmp1);
}

/*
 * Run iterations i_start .. i_start+n_iters-1 of the outer loop (solverwrapper.jl:23-29),
 * stopping when the status leaves :Continue.  History records are appended to hist
 * (REC_LEN doubles each).  If trace != NULL the iterate x after each executed iteration is
 * stored (row k = iteration i_start+k).  Returns the number of iterations executed.
 */
FOSOR_API int64_t fosor_run(void *h, int64_t i_start, int64_t n_iters, int64_t checki, double eps, double *hist,
                            int64_t hist_cap, int64_t *hist_len, double *trace, int32_t *status_out)
{
    model_t *M = (model_t *)h;
    M->checki = checki;
    M->eps = eps;
    M->hist = hist; M->hist_cap = hist_cap; M->hist_len = 0;
    int64_t done = 0;
    for (int64_t i = i_start; i < i_start + n_iters; i++) {
        M->cur_i = i;       /* solverwrapper.jl:24 */
        do_step(M);         /* :25 */
        if (trace) memcpy(trace + done * M->N, M->x, sizeof(double) * (size_t)M->N);
        done += 1;
        if (M->status != ST_CONTINUE) break; /* :26-28 */
    }
    if (hist_len) *hist_len = M->hist_len;
    if (status_out) *status_out = M->status;
    M->hist = NULL;
    return done;
}

/*
 * The tail of iterate() (solverwrapper.jl:31-34): guess = getsol(...); forced check when the
 * last iteration was not a check iteration.  One record may be appended to hist.
 */
FOSOR_API void fosor_finish(void *h, double *guess, double *hist, int64_t *hist_len, int32_t *status_out)
{
    model_t *M = (model_t *)h;
    M->hist = hist; M->hist_cap = 1; M->hist_len = 0;
    do_getsol(M, guess);
    if (!M->checked) checkstatus(M, guess, 1);
    if (hist_len) *hist_len = M->hist_len;
    if (status_out) *status_out = M->status;
    M->hist = NULL;
}

/* solve!(model) (solverwrapper.jl:2-17) = run from i=1 + finish; status reset like a fresh status_generator call */
FOSOR_API int64_t fosor_solve(void *h, int64_t max_iters, int64_t checki, double eps, double *guess, double *hist,
                              int64_t hist_cap, int64_t *hist_len, int32_t *status_out)
{
    model_t *M = (model_t *)h;
    M->status = ST_CONTINUE;
    M->checked = 0;
    for (int64_t i = 0; i < M->N; i++) M->prev[i] = NAN;
    int64_t hl = 0, hl2 = 0;
    int32_t st = 0;
    int64_t done = fosor_run(h, 1, max_iters, checki, eps, hist, hist_cap, &hl, NULL, &st);
    double *tail = (hist && hl < hist_cap) ? hist + hl * REC_LEN : NULL;
    fosor_finish(h, guess, tail, &hl2, &st);
    if (!tail) hl2 = M->hist_len;
    if (hist_len) *hist_len = hl + hl2;
    if (status_out) *status_out = st;
    return done;
}

/* HSDE_populatesolution (HSDE.jl:49-61): x/tau, y/tau, s/tau */
FOSOR_API void fosor_populate_solution(void *h, const double *guess, double *x, double *y, double *s)
{
    model_t *M = (model_t *)h;
    int64_t m = M->A->m, n = M->A->n, l = m + n + 1;
    double tau = guess[l - 1];
    for (int64_t j = 0; j < n; j++) x[j] = guess[j] / tau;
    for (int64_t i = 0; i < m; i++) y[i] = guess[n + i] / tau;
    for (int64_t i = 0; i < m; i++) s[i] = guess[l + n + i] / tau;
}

/* ------------------------------------------------------------------------------------
 * Unit-level entry points (the restated unit tests of SURVEY.md section 4 drive these)
 * ---------------------------------------------------------------------------------- */
FOSOR_API void fosor_q_mul(void *h, const double *B, double *Y, int32_t transpose)
{
    model_t *M = (model_t *)h;
    if (transpose) linop_mul_t(Y, &M->S1->op, B);
    else linop_mul(Y, &M->S1->op, B);
}
FOSOR_API void fosor_kkt_mul(void *h, const double *x, double *y) { kkt_mul(y, &((model_t *)h)->S1->op, x); }
FOSOR_API void fosor_affine_prox(void *h, const double *x, double *y) { prox_S1((model_t *)h, y, x); }
FOSOR_API void fosor_cone_prox(void *h, const double *x, double *y) { prox_S2((model_t *)h, y, x); }
FOSOR_API void fosor_a_mul(void *h, const double *x, double *y, int32_t transpose)
{
    model_t *M = (model_t *)h;
    if (transpose) csc_mul_t(y, M->A, x);
    else csc_mul(y, M->A, x);
}

/* HSDEMatrix.prox! (a14) on a fresh CG state */
FOSOR_API void fosor_hsdematrix_prox(void *h, const double *x, double *y)
{
    model_t *M = (model_t *)h;
    hsdem_t H;
    int64_t N = M->N;
    H.Q = M->S1->op;
    H.r = (double *)calloc((size_t)N, sizeof(double));
    H.p = (double *)calloc((size_t)N, sizeof(double));
    H.z = (double *)calloc((size_t)N, sizeof(double));
    H.xinit = (double *)calloc((size_t)N, sizeof(double));
    H.firstrun = 1;
    hsdem_prox(y, &H, x);
    (void)hsdem_mul;
    free(H.r); free(H.p); free(H.z); free(H.xinit);
}

/* conjugategradient!(x, A, b, r, p, Ap; tol, max_iters) on a CSC matrix (test/conjugateGradient.jl) */
FOSOR_API int64_t fosor_cg_csc(int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                               int64_t index_base, const double *b, double *x, double tol, int64_t max_iters)
{
    csc_t *A = csc_new(n, n, colptr, rowval, nzval, index_base);
    double *r = (double *)calloc((size_t)n, sizeof(double));
    double *p = (double *)calloc((size_t)n, sizeof(double));
    double *Ap = (double *)calloc((size_t)n, sizeof(double));
    if (tol < 0) tol = (double)n * JULIA_EPS; /* default tol = size(A,2)*eps() :31 */
    int64_t it = conjgrad(x, mul_csc_cb, A, b, r, p, Ap, n, tol, max_iters);
    free(r); free(p); free(Ap);
    csc_free(A);
    return it;
}

/* stand-alone cone projections for unit tests: dual != 0 -> proxDual! */
FOSOR_API int32_t fosor_prox_cone(int32_t type, int32_t dual, const double *x, double *y, int64_t len)
{
    return dual ? prox_cone_dual(y, type, x, len) : prox_cone(y, type, x, len);
}
