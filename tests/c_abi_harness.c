/*
 * c_abi_harness.c -- plain-C driver of the drop-in boundary: #include "fos_b200.h", nothing else of the product.
 *
 * TEST INFRASTRUCTURE.  What ctypes cannot show -- that the header itself (not hand-typed signatures) compiles as
 * C, links against libfos_b200.so, and drives a whole solve -- is shown here.  It is what the Julia `ccall` glue
 * does, written in the one compiled host language this image has.
 *
 *   c_abi_harness --no-gpu   header + link + "no CPU fallback" check (runs on the CPU-only build box)
 *   c_abi_harness            config 1 end to end on the GPU: README NNLS 40 x 50 in conic form
 *                            (m = 91, n = 51, K1 = SOC(41) + NonNeg(50)), GAP(0.5, 2.0, 2.0, max_iters = 2000)
 *                            = solve!(model) of solverwrapper.jl:2-17 through fos_create / fos_load_conic_csc /
 *                            fos_set_algorithm / fos_solve, checked against the CPU oracle (libfos_oracle.so:
 *                            same status, same iteration count, same check iterations, solution within 1e-4) and
 *                            against the problem itself (x >= 0, reported p/d residuals reproduced in C).
 *
 * gcc -std=c11 -I include tests/c_abi_harness.c -L firstordersolvers.jl_b200 -lfos_b200 -L oracle -lfos_oracle -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fos_b200.h"

/* the oracle's C entry points used here (oracle/fos_oracle.c) */
void *fosor_create_conic(int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                         int64_t index_base, const double *b, const double *c, int64_t nc1, const int32_t *types1,
                         const int64_t *lens1, int64_t nc2, const int32_t *types2, const int64_t *lens2);
void fosor_set_algorithm(void *h, int32_t alg, double alpha, double alpha1, double alpha2, double beta, int64_t iproj);
int64_t fosor_solve(void *h, int64_t max_iters, int64_t checki, double eps, double *guess, double *hist,
                    int64_t hist_cap, int64_t *hist_len, int32_t *status_out);
void fosor_destroy(void *h);

static uint64_t rng_state = 0x853c49e6748fea9bULL;
static double urand(void)
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return ((double)(rng_state >> 11) + 0.5) / 9007199254740992.0;
}
static double nrand(void) { return sqrt(-2.0 * log(urand())) * cos(6.283185307179586 * urand()); }

#define CHECK(cond, ...)                      \
    do {                                      \
        if (!(cond)) {                        \
            fprintf(stderr, "FAIL: ");        \
            fprintf(stderr, __VA_ARGS__);     \
            fprintf(stderr, "\n");            \
            return 1;                         \
        }                                     \
    } while (0)

int main(int argc, char **argv)
{
    const int no_gpu = argc > 1 && strcmp(argv[1], "--no-gpu") == 0;
    CHECK(fos_abi_version() == FOS_ABI_VERSION, "ABI version %d, header says %d", fos_abi_version(), FOS_ABI_VERSION);
    fos_handle_t h = NULL;
    int32_t rc = fos_create(&h, 0);
    if (no_gpu) {
        if (rc == FOS_OK) { /* a GPU is present after all: fine */
            fos_destroy(h);
            printf("ok (device present)\n");
            return 0;
        }
        CHECK(rc == FOS_ERR_CUDA && h == NULL, "without a GPU fos_create must fail with FOS_ERR_CUDA, got %d", rc);
        CHECK(strlen(fos_last_error(NULL)) > 0, "no error text");
        printf("ok (no GPU: %s)\n", fos_last_error(NULL));
        return 0;
    }
    CHECK(rc == FOS_OK, "fos_create: %s", fos_last_error(NULL));

    /* ---- config 1: minimise ||D x - d|| s.t. x >= 0  as  min t  s.t. (t, D x - d) in SOC, x in NonNeg ---- */
    enum { R = 40, Cc = 50, M = R + 1 + Cc, N = Cc + 1 };
    static double D[R][Cc], d[R];
    for (int i = 0; i < R; i++)
        for (int j = 0; j < Cc; j++) D[i][j] = nrand();
    for (int i = 0; i < R; i++) d[i] = nrand();
    /* A = [-1 0; 0 -D; 0 -I] in CSC, 0-based here (Julia passes its 1-based arrays with index_base = 1) */
    static int64_t colptr[N + 1], rowval[1 + Cc * (R + 1)];
    static double nzval[1 + Cc * (R + 1)];
    int64_t nnz = 0;
    colptr[0] = 0;
    rowval[nnz] = 0; nzval[nnz++] = -1.0;
    colptr[1] = nnz;
    for (int j = 0; j < Cc; j++) {
        for (int i = 0; i < R; i++) { rowval[nnz] = 1 + i; nzval[nnz++] = -D[i][j]; }
        rowval[nnz] = 1 + R + j; nzval[nnz++] = -1.0;
        colptr[2 + j] = nnz;
    }
    static double b[M], c[N];
    for (int i = 0; i < R; i++) b[1 + i] = -d[i];
    c[0] = 1.0;
    const int32_t t1[2] = {FOS_CONE_SOC, FOS_CONE_NONNEG}, t2[1] = {FOS_CONE_FREE};
    const int64_t l1[2] = {R + 1, Cc}, l2[1] = {N};

    rc = fos_load_conic_csc(h, M, N, colptr, rowval, nzval, 0, b, c, 2, t1, l1, 1, t2, l2, FOS_STORAGE_AUTO);
    CHECK(rc == FOS_OK, "fos_load_conic_csc: %s", fos_last_error(h));
    rc = fos_set_algorithm(h, FOS_ALG_GAP, 0.5, 2.0, 2.0, 0.0, 100); /* GAP(0.5, 2.0, 2.0) README.md:27 */
    CHECK(rc == FOS_OK, "fos_set_algorithm: %s", fos_last_error(h));
    const int64_t L = M + N + 1, LEN = 2 * L;
    CHECK(fos_iterate_length(h) == LEN, "iterate length %lld", (long long)fos_iterate_length(h));
    CHECK(fos_set_initial_iterate(h) == FOS_OK, "fos_set_initial_iterate");

    enum { CAP = 32 };
    static double guess[2 * (M + N + 1)], rec[CAP * FOS_REC_LEN];
    int64_t done = 0, nrec = 0;
    int32_t st = -1;
    rc = fos_solve(h, 2000, 100, 1e-5, guess, LEN, &done, &st, rec, CAP, &nrec);
    CHECK(rc == FOS_OK, "fos_solve: %s", fos_last_error(h));

    /* ---- the same model on the CPU oracle ---- */
    void *o = fosor_create_conic(M, N, colptr, rowval, nzval, 0, b, c, 2, t1, l1, 1, t2, l2);
    CHECK(o != NULL, "oracle rejected the cones");
    fosor_set_algorithm(o, 0, 0.5, 2.0, 2.0, 0.0, 100);
    static double oguess[2 * (M + N + 1)], ohist[CAP * FOS_REC_LEN];
    int64_t ohl = 0;
    int32_t ost = -1;
    const int64_t odone = fosor_solve(o, 2000, 100, 1e-5, oguess, ohist, CAP, &ohl, &ost);
    fosor_destroy(o);
    if (ost == FOS_STATUS_CONTINUE) ost = FOS_STATUS_INDETERMINATE; /* HSDE.jl:56-59 */

    printf("GPU:    status %d after %lld iterations, %lld check records\n", st, (long long)done, (long long)nrec);
    printf("oracle: status %d after %lld iterations, %lld check records\n", ost, (long long)odone, (long long)ohl);
    CHECK(st == ost, "status differs");
    CHECK(done == odone, "iteration count differs");
    CHECK(nrec == ohl, "number of status checks differs");
    for (int64_t k = 0; k < nrec && k < CAP; k++)
        CHECK(rec[k * FOS_REC_LEN] == ohist[k * FOS_REC_LEN] && rec[k * FOS_REC_LEN + 9] == ohist[k * FOS_REC_LEN + 9],
              "check %lld: iteration / status differ", (long long)k);
    /* solution (HSDE_populatesolution, HSDE.jl:49-61): x/tau */
    const double tau = guess[L - 1], otau = oguess[L - 1];
    CHECK(tau > 0.0 && otau > 0.0, "tau = %g / %g", tau, otau);
    double worst = 0.0, scale = 0.0, minx = 1e300;
    for (int j = 0; j < N; j++) {
        const double xg = guess[j] / tau, xo = oguess[j] / otau;
        if (fabs(xg - xo) > worst) worst = fabs(xg - xo);
        if (fabs(xo) > scale) scale = fabs(xo);
        if (j > 0 && xg < minx) minx = xg;
    }
    printf("solution: max |x_gpu - x_oracle| = %.3e (scale %.3e), min x = %.3e, t = %.6f\n", worst, scale, minx,
           guess[0] / tau);
    CHECK(worst <= 1e-4 * scale, "solutions differ by %.3e", worst);
    CHECK(minx > -1e-4, "x >= 0 violated: %.3e", minx);
    /* the objective t must equal ||D x - d|| at the returned point to the solver's accuracy */
    double res = 0.0;
    for (int i = 0; i < R; i++) {
        double s = -d[i];
        for (int j = 0; j < Cc; j++) s += D[i][j] * guess[1 + j] / tau;
        res += s * s;
    }
    res = sqrt(res);
    printf("||D x - d|| = %.6f\n", res);
    CHECK(fabs(res - guess[0] / tau) <= 1e-3 * (1.0 + res), "objective %.6f vs residual norm %.6f", guess[0] / tau, res);
    /* the last record's primal residual, recomputed here: p = ||A x/tau + s/tau - b|| / (1 + ||b||)  (HSDEStatus.jl:34);
     * only the forced final check (solverwrapper.jl:32-34) is evaluated on the returned point */
    if (done % 100 != 0) {
        double nb = 0.0, acc = 0.0;
        static double ax[M];
        for (int i = 0; i < M; i++) { ax[i] = 0.0; nb += b[i] * b[i]; }
        for (int j = 0; j < N; j++)
            for (int64_t k = colptr[j]; k < colptr[j + 1]; k++) ax[rowval[k]] += nzval[k] * guess[j];
        for (int i = 0; i < M; i++) {
            const double v = ax[i] / tau + guess[L + N + i] / tau - b[i];
            acc += v * v;
        }
        const double p_here = sqrt(acc) / (1.0 + sqrt(nb));
        const double p_rec = rec[(nrec - 1) * FOS_REC_LEN + 1];
        printf("final check: p = %.6e (recomputed %.6e)\n", p_rec, p_here);
        CHECK(fabs(p_here - p_rec) <= 1e-8 + 1e-6 * p_rec, "recorded primal residual not reproduced");
    }
    CHECK(fos_destroy(h) == FOS_OK, "fos_destroy");
    printf("ok\n");
    return 0;
}
