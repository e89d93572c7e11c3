/*
 * fos_b200.h -- C ABI of the B200-native hot path of FirstOrderSolvers.jl.
 *
 * Plain C, opaque handle, caller-owned host buffers, no exceptions across the boundary.
 * Every function returns an int32 status: 0 = ok, <0 = error (see FOS_ERR_*); the text of
 * the last error is available from fos_last_error().  A handle is NOT thread-safe (the
 * reference is single-threaded and stateful per model: CG warm start, S.i, GAPA's alpha12,
 * FISTA's t); distinct handles are independent.
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * /root/reference/src).  The Julia-side binding (ccall) is shown in INTEGRATION.md and
 * shipped in firstordersolvers.jl_b200/julia/FirstOrderSolversB200.jl.
 *
 * Vector convention (identical to the reference's): for the HSDE conic form the iterate is
 * z = [x(n); y(m); tau; r(n); s(m); kappa], length 2(m+n+1)
 * (problemforms/HSDE/HSDEStatus.jl:93-102, cones.jl:125-134); for the affine/feasibility
 * form it is [x(an); z(am)] (utilities/affinepluslinear.jl:85-90).  All data are FP64.
 */
#ifndef FOS_B200_H
#define FOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define FOS_ABI_VERSION 1

typedef struct fos_handle_s *fos_handle_t;

/* ---- error codes --------------------------------------------------------------------- */
#define FOS_OK 0
#define FOS_ERR_INVALID (-1)     /* bad argument / bad state                                */
#define FOS_ERR_CUDA (-2)        /* CUDA runtime or driver failure (no CPU fallback exists) */
#define FOS_ERR_UNSUPPORTED (-3) /* e.g. cone type outside the hot-path scope               */
#define FOS_ERR_NOMEM (-4)
#define FOS_ERR_COMM (-5)        /* NCCL failure                                            */

/* ---- cone codes: cones.jl:4-14 (conemap) --------------------------------------------- */
#define FOS_CONE_FREE 0
#define FOS_CONE_ZERO 1
#define FOS_CONE_NONNEG 2
#define FOS_CONE_NONPOS 3
#define FOS_CONE_SOC 4
#define FOS_CONE_SOCROT 5    /* IndRotatedSOC on (x1, x2, w): 2 x1 x2 >= ||w||^2, x1, x2 >= 0 */
#define FOS_CONE_SDP 6
#define FOS_CONE_EXPPRIMAL 7 /* IndExpPrimal: cl{(r,s,t): s > 0, s exp(r/s) <= t}, triples */
#define FOS_CONE_EXPDUAL 8   /* IndExpDual (its dual cone) */

/* ---- algorithm codes: solvers/{gap,gapa,fista,dykstra,gapproj}.jl --------------------- */
#define FOS_ALG_GAP 0     /* GAP(alpha,alpha1,alpha2) gap.jl:6-13; DR/AP are GAP(a,2,2)/GAP(a,1,1) solvers.jl:10-11 */
#define FOS_ALG_GAPA 1    /* GAPA(alpha,beta)          gapa.jl:9-15   */
#define FOS_ALG_FISTA 2   /* FISTA(alpha)              fista.jl:6-11  */
#define FOS_ALG_DYKSTRA 3 /* Dykstra()                 dykstra.jl:6-10 */
#define FOS_ALG_GAPP 4    /* GAPP(alpha,alpha1,alpha2; iproj) gapproj.jl:6-14 */

/* ---- status codes: HSDEStatus.jl:53-63, HSDE.jl:56-59 -------------------------------- */
#define FOS_STATUS_CONTINUE 0
#define FOS_STATUS_OPTIMAL 1
#define FOS_STATUS_UNBOUNDED 2
#define FOS_STATUS_INFEASIBLE 3
#define FOS_STATUS_INDETERMINATE 4

/* ---- status-check record: one per executed convergence check -------------------------
 * HSDE form  (HSDEStatus.jl:125-139 savedata + :46 cgiter):
 *     [0] i  [1] p  [2] d  [3] g  [4] ctx  [5] bty  [6] kappa  [7] tau  [8] cgiter  [9] status
 * Feasibility form (FeasibilityStatus.jl:94-103):
 *     [0] i  [1] err  [2..7] 0  [8] cgiter  [9] status                                    */
#define FOS_REC_LEN 10

/* ---- matrix storage selector for the loaders ------------------------------------------ */
#define FOS_STORAGE_AUTO 0   /* dense when nnz/(m*n) > 0.25, else sparse */
#define FOS_STORAGE_DENSE 1  /* row-major FP64 tiles streamed by TMA (K1) */
#define FOS_STORAGE_SPARSE 2 /* CSR + CSC copies, int32 indices          */

/* ---- where a matrix argument lives ---------------------------------------------------- */
#define FOS_MEM_HOST 0
#define FOS_MEM_DEVICE 1 /* device pointer on the handle's device (e.g. a torch tensor's data_ptr) */

/* ====================================================================================== */
/* lifecycle                                                                              */
/* ====================================================================================== */
int32_t fos_abi_version(void);
/* Creates a solver handle bound to CUDA device `device`.  Fails (FOS_ERR_CUDA) when no
 * sm_100 device is present: there is no CPU path.  Replaces the GAPData/GAPAData/... structs
 * built by init_algorithm! (solvers/gap.jl:23-28 etc.). */
int32_t fos_create(fos_handle_t *out, int32_t device);
int32_t fos_destroy(fos_handle_t h);
/* Text of the last error on `h` (or of the last failed fos_create when h == NULL). */
const char *fos_last_error(fos_handle_t h);
/* Tunables / debug knobs (all optional).  Keys:
 *   "matvec_impl"  0 = TMA-staged fused kernel (default), 1 = plain two-kernel reference path
 *   "grid_ctas"    number of persistent CTAs of the fused kernel (default = #SMs)
 *   "cg_batch"     CG iterations enqueued per host synchronisation (default adaptive = 0)
 *   "fuse_rhs"     1 (default) = fold the right-hand-side product of the affine projection into the
 *                  initial CG residual (k+1 passes over A per projection instead of k+2; same
 *                  mathematics, sums associated differently); 0 = build rhs in the reference's order
 *   "fuse_tail"    1 (default) = conic form: everything of a CG iteration after the pass over A (peer
 *                  exchange, KKT epilogue, both dot products, x/r/p updates, stop test) runs in ONE
 *                  cooperative kernel; 0 = one kernel per step (K2, K3 update, K3 direction)
 *   "psd_warm"     1 (default) = large PSD cones start their Jacobi sweeps from the eigenvector basis of the
 *                  previous projection of the same cone (consecutive iterates are close: 2-4 sweeps instead
 *                  of ~10); 0 = always start from the identity.  Set after loading.
 *   "exchange_impl" multi-GPU: 1 = fused peer-memory exchange (after fos_comm_p2p_import), 0 = NCCL
 *   "profile_matvec" 1 = CUDA events around every mat-vec launch (read back with fos_get_info)
 *   "hybrid_rows"  single-problem dense loads classify the rows on the device and keep only the contiguous block that
 *                  holds every row with more than n/8 non-zeros as dense tiles; the other non-empty rows (the -I blocks
 *                  of SOC / NonNeg constraints on the variables) are stored as CSR + CSC.  2 (default) = only when that
 *                  saves at least 64 MB per pass (config 3: 19.2 -> 16.0 GB, +17 % iterations/s); 1 = whenever it saves
 *                  2 % of the bytes; 0 = never.  Set before loading; single rank only.
 *   "batch_hybrid" 1 (default) = batch mode keeps rows with <= n/8 non-zeros out of the dense tiles (CSR + CSC) and
 *                  skips empty rows; 0 = every row is streamed as dense FP64.  Set before loading the batch.
 *   "batch_ctas"   persistent CTAs of the batch kernel (default = #SMs)
 *   "k1_balance"   the fused mat-vec runs one persistent CTA per SM and ncu shows the SMs finishing between 0.85 and 0.99 of
 *                  a pass.  1 / 2 = at load, time a few passes per SM, bind the work ranges to SMs and size them to the
 *                  measured speeds (1: only with >= 4 row groups per SM; cached per device and shard shape for the life of
 *                  the process, so equal problems get equal plans and bit-identical results).  Measured on B200
 *                  (profiles/r2_k1_balance.md): the spread of the per-SM times shrinks but the pass does not get faster --
 *                  an SM that finishes early hands its share of the HBM bandwidth to the others; the pass is bound by
 *                  HBM, not by the slowest SM.  Default 0 (even split).  Set before loading.
 *   "psd_warp_max_d" PSD cones up to this order (default and at most 16) are projected by one warp each, eight cones per
 *                  CTA; up to 48 by 128 threads, up to 112 by one 512-thread CTA, larger ones by the cooperative kernel.
 *                  Process-wide; applies to problems loaded afterwards.
 *   "tail_trace"   1 = the fused CG tail records the SM cycles of each of its phases (fos_get_tail_trace)
 *   "use_graphs"   reserved                                                                */
int32_t fos_set_option(fos_handle_t h, const char *key, double value);

/* ====================================================================================== */
/* multi-GPU: one process per GPU, A row-sharded, NCCL all-reduce of the A' partial sums  */
/* (new functionality; the reference has no parallelism -- SURVEY.md 8e)                  */
/* ====================================================================================== */
#define FOS_COMM_ID_BYTES 128
int32_t fos_comm_unique_id(uint8_t *id_out /* FOS_COMM_ID_BYTES */);
/* An all-zero id creates no NCCL communicator: the handle then relies on the peer-memory exchange below
 * (fos_comm_p2p_export / _import must follow the load); used where NCCL cannot run, e.g. several ranks on one GPU. */
int32_t fos_comm_init(fos_handle_t h, int32_t rank, int32_t nranks, const uint8_t *id /* FOS_COMM_ID_BYTES */);
/* Fused exchange over NVLink peer memory (optional, after fos_load_conic_dense on every rank): replaces
 * the fold kernel + ncclAllReduce of every pass over A by ONE kernel that publishes this rank's partial
 * sums in a CUDA-IPC-shared slot and gathers the peers' slots with direct loads over NVLink, summing
 * in rank order (bitwise identical results on every rank).  Each rank exports a 64-byte handle, the
 * caller all-gathers them (rank order) and hands the table to every rank.  One process per GPU. */
#define FOS_IPC_HANDLE_BYTES 64
int32_t fos_comm_p2p_export(fos_handle_t h, uint8_t *handle_out /* FOS_IPC_HANDLE_BYTES */);
int32_t fos_comm_p2p_import(fos_handle_t h, const uint8_t *handles /* nranks * FOS_IPC_HANDLE_BYTES */);

/* ====================================================================================== */
/* problem loading -- replaces loadproblem! (FOSSolverInterface.jl:27-64) + HSDE()        */
/* (problemforms/HSDE/HSDE.jl:7-29, indirect branch)                                      */
/* ====================================================================================== */
/* Conic form: minimise c'x  s.t.  b - A x in K1,  x in K2  (MathProgBase convention).
 * A is Julia's SparseMatrixCSC{Float64,Int64} passed as is (index_base = 1) or 0-based.
 * Cones are contiguous and ordered (cones.jl:66-72): cone k covers len[k] consecutive
 * entries; sum(len1) == m, sum(len2) == n.  The library copies everything to the device. */
int32_t fos_load_conic_csc(fos_handle_t h, int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval,
                           const double *nzval, int64_t index_base, const double *b, const double *c,
                           int64_t ncones1, const int32_t *cone_type1, const int64_t *cone_len1, int64_t ncones2,
                           const int32_t *cone_type2, const int64_t *cone_len2, int32_t storage);

/* Same model, A given dense row-major (leading dimension lda >= n, in elements), on the host
 * or already on the device (benchmark shapes; FOSSolverInterface.jl:27-29 is the reference's
 * dense entry).  With row sharding (after fos_comm_init) `A` holds only rows
 * [row_begin, row_begin+row_count) of the m x n matrix; b, c and the cones are global.
 * Without sharding pass row_begin = 0, row_count = m. */
int32_t fos_load_conic_dense(fos_handle_t h, int64_t m, int64_t n, const double *A, int64_t lda, int32_t a_location,
                             int64_t row_begin, int64_t row_count, const double *b, const double *c,
                             int64_t ncones1, const int32_t *cone_type1, const int64_t *cone_len1, int64_t ncones2,
                             const int32_t *cone_type2, const int64_t *cone_len2);

/* Affine/feasibility form (problemforms/Feasibility/Feasibility.jl:2-6 with
 * S1 = AffinePlusLinear(A, b, q, beta; decreasing_accuracy) utilities/affinepluslinear.jl:71-79
 * and S2 = a ConeProduct over the an+am entries of [x; z]).  Status = FeasibilityStatus. */
int32_t fos_load_affine_csc(fos_handle_t h, int64_t am, int64_t an, const int64_t *colptr, const int64_t *rowval,
                            const double *nzval, int64_t index_base, const double *b, const double *q, int32_t beta,
                            int32_t decreasing_accuracy, int64_t ncones, const int32_t *cone_type,
                            const int64_t *cone_len, int32_t storage);

/* ====================================================================================== */
/* algorithm and iterate                                                                  */
/* ====================================================================================== */
/* Constructors of solvers/{gap,gapa,fista,dykstra,gapproj}.jl (a20).  Resets the algorithm data (alpha12 = 2, t = 1,
 * p = q = 0) like init_algorithm!; does NOT reset S1's warm start / call counter. */
int32_t fos_set_algorithm(fos_handle_t h, int32_t alg, double alpha, double alpha1, double alpha2, double beta,
                          int64_t iproj);
/* IndBox(lo, hi) (ProximalOperators; S2 of test/testfeasibility.jl:10) on entries [start, start+len) (0-based) of
 * the Feasibility iterate [x; z]: the projection clamps to [lo, hi] (use +-INFINITY for one-sided boxes).  Call
 * after fos_load_affine_csc; it overrides the elementwise cone given there for that range. */
int32_t fos_set_box(fos_handle_t h, int64_t start, int64_t len, double lo, double hi);
/* LineSearchWrapper(alg; lsinterval) (wrappers/linesearch.jl:19-75) around GAP / GAPA (the algorithms with
 * support_linesearch, gap.jl:89, gapa.jl:117): on iterations i % lsinterval == 0 the step is replaced by one
 * relaxed S1/S2 pass (with the status check inside S2!), 31 trial steps x = x0 + alpha*res, alpha = 0.1*1.8^(k+1),
 * each scored by ||x - S2!(S1!(x))||, and x = x0 + alpha_best*res.  alpha_best: fos_get_info(h, 8).  The
 * reference's 33 println lines per search are not reproduced.  lsinterval = 0 removes the wrapper; other
 * algorithms ignore it (the reference logs an error at construction). */
int32_t fos_set_linesearch(fos_handle_t h, int64_t lsinterval);
/* direct = true (the `direct` field of every algorithm, HSDE.jl:10-15; the default of GAPP, gapproj.jl:14):
 * S1 becomes IndAffine([Q -I], 0), the exact projection, instead of the truncated CG solve.  The one-time
 * work (W = (I + Q Q')^-1, dense, on the device) happens in this call; every projection is then two passes
 * over A and one over W.  Conic form only, l = m+n+1 up to 16384, no row sharding.  on = 0 returns to CG. */
int32_t fos_set_direct(fos_handle_t h, int32_t on);
int64_t fos_iterate_length(fos_handle_t h); /* 2(m+n+1) or an+am; <0 on error */
/* initx / getinitialvalue (solverwrapper.jl:10, HSDE.jl:40-47, Feasibility.jl:57-58). */
int32_t fos_set_iterate(fos_handle_t h, const double *z, int64_t len);
int32_t fos_set_initial_iterate(fos_handle_t h); /* z0 = 0 except tau = kappa = 1 (HSDE) / zeros */
int32_t fos_get_iterate(fos_handle_t h, double *z, int64_t len);
/* Internal vectors for debug=2, parity tests and checkpointing.  which: 0 x, 1 tmp1, 2 tmp2
 * (relaxed), 3 S1.cgdata.xinit (CG warm start), 4 S1.rhs, 5 last unrelaxed S2 projection (what
 * checkstatus sees), 6 FISTA y (fista.jl:15), 7 Dykstra p, 8 Dykstra q (dykstra.jl:13-14). */
int32_t fos_get_state(fos_handle_t h, int32_t which, double *buf, int64_t len);
/* Restores one of the persistent vectors (which = 0, 3, 6, 7 or 8).  Setting 3 also clears S1's
 * first-run flag (affinepluslinear.jl:101-104).  The reference cannot restore its CG warm start
 * (checkpoint/resume); this is a superset used by the lock-step parity tests. */
int32_t fos_set_state(fos_handle_t h, int32_t which, const double *buf, int64_t len);
/* Scalars.  which: 0 S1.i (call counter, affinepluslinear.jl:66), 1 S1.cgiter (:67),
 * 2 GAPA alpha12 (gapa.jl:18), 3 FISTA t (fista.jl:14), 4 CG-hit-max-iters flag (the @warn of
 * conjugategradients.jl:53), 5 total CG iterations so far, 6 total passes over A so far,
 * 7 kernel launches so far, 8 GAPP alpha_best of the last projected step; with the option
 * "profile_matvec" = 1: 9 / 10 summed milliseconds / count of 2-RHS mat-vec launches, 11 / 12 the
 * same for 1-RHS launches, 13 predicated no-op launches; 14 algorithmic bytes of one pass over A,
 * 15 number of SMs, 16 / 17 summed milliseconds / count of executed fused CG-tail launches,
 * 18 storage of A (1 dense, 2 sparse, 3 hybrid), 19 / 20 rows of the dense block / rows kept sparse
 * under hybrid row storage (0 otherwise), 21 = 1 when the work ranges of the fused mat-vec are sized to the SMs'
 * measured speed ("k1_balance"), 22 / 23 (max - min) / mean of the per-SM times of a pass before / after balancing. */
int32_t fos_get_info(fos_handle_t h, int32_t which, double *out);
/* Restores a scalar: which = 0 (S1.i), 2 (alpha12) or 3 (FISTA t). */
int32_t fos_set_info(fos_handle_t h, int32_t which, double value);

/* ====================================================================================== */
/* the hot loop -- replaces iterate() (solverwrapper.jl:20-41) and step() of every solver */
/* ====================================================================================== */
/* A fresh status object, as model.status_generator builds one per solve! (solverwrapper.jl:13,
 * HSDE.jl:26-27, Feasibility.jl:78-79): status = Continue, checked = false, prev = NaN.  Does not
 * touch the iterate, S1's warm start / call counter or the algorithm data (SURVEY a-Q 2). */
int32_t fos_begin_solve(fos_handle_t h);
/* Runs iterations i = i_start .. i_start+n_iters-1 (solverwrapper.jl:23-29), stopping early
 * when a check sets a status other than Continue.  Every check (i % checki == 0) appends one
 * record (FOS_REC_LEN doubles) to `records` (capacity rec_cap records; extra checks are
 * counted in *n_rec but not stored).  If trace != NULL it receives the iterate after each
 * executed iteration (row k = iteration i_start+k, row length = fos_iterate_length). */
int32_t fos_run(fos_handle_t h, int64_t i_start, int64_t n_iters, int64_t checki, double eps, int64_t *iters_done,
                int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec, double *trace);
/* Tail of iterate() (solverwrapper.jl:31-34): guess = getsol(alg,data,x) = P2(P1(x)) -- which
 * runs one more CG solve and advances S1.i (gap.jl:82-87) -- then the forced final check when
 * the last iteration was not a check iteration.  At most one record is written. */
int32_t fos_finish(fos_handle_t h, double *guess, int64_t len, double *record, int64_t *n_rec, int32_t *status);
/* solve!(model) (solverwrapper.jl:2-17) on the current iterate: fos_begin_solve, run from
 * i = 1, finish.  Status Continue is reported as Indeterminate (HSDE.jl:56-59). */
int32_t fos_solve(fos_handle_t h, int64_t max_iters, int64_t checki, double eps, double *guess, int64_t len,
                  int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec);

/* ====================================================================================== */
/* batch mode: B independent conic problems of identical shape (same m, n, cone layout;  */
/* different A, b, c), one persistent CTA per problem, no host round trips (new; the     */
/* reference solves one model at a time -- SURVEY.md 8b "fos_load_conic_batch /         */
/* fos_run_batch", 8e batch split).  Every per-problem array has a leading batch          */
/* dimension.  The algorithm is chosen with fos_set_algorithm (GAP/DR/AP, GAPA, FISTA,    */
/* Dykstra); a handle is either in batch mode or in single-problem mode.                  */
/* ====================================================================================== */
/* A: B matrices, m x n row-major, leading dimension lda, problem stride pstride (elements), on the host
 * or on the device; b: B x m, c: B x n (host).  Cones as in fos_load_conic_csc (SDP not offered). */
int32_t fos_load_conic_dense_batch(fos_handle_t h, int64_t nprob, int64_t m, int64_t n, const double *A, int64_t lda,
                                   int64_t pstride, int32_t a_location, const double *b, const double *c,
                                   int64_t ncones1, const int32_t *cone_type1, const int64_t *cone_len1,
                                   int64_t ncones2, const int32_t *cone_type2, const int64_t *cone_len2);
int64_t fos_batch_size(fos_handle_t h); /* number of problems; <0 when the handle is not in batch mode */
/* z: B x fos_iterate_length, or NULL for the initial value of every problem (HSDE.jl:40-47). */
int32_t fos_set_iterate_batch(fos_handle_t h, const double *z);
int32_t fos_get_iterate_batch(fos_handle_t h, double *z);
/* Same selectors as fos_get_state / fos_set_state (0 x, 1 tmp1, 2 tmp2, 3 xinit, 5 last unrelaxed S2
 * projection, 6 FISTA y, 7/8 Dykstra p/q); buf: B x fos_iterate_length. */
int32_t fos_get_state_batch(fos_handle_t h, int32_t which, double *buf);
int32_t fos_set_state_batch(fos_handle_t h, int32_t which, const double *buf);
/* Per-problem scalars, B values.  which: 0 S1.i, 1 S1.cgiter, 2 alpha12, 3 FISTA t (settable: 0, 2, 3);
 * read-only: 5 total CG iterations, 6 total passes over A. */
int32_t fos_get_info_batch(fos_handle_t h, int32_t which, double *out);
int32_t fos_set_info_batch(fos_handle_t h, int32_t which, const double *values);
int32_t fos_begin_solve_batch(fos_handle_t h);
/* fos_run for every problem: iterations i_start .. i_start+n_iters-1, each problem stopping on its own
 * when a check sets its status.  iters_done, status, n_rec: B entries; records: B x rec_cap x FOS_REC_LEN. */
int32_t fos_run_batch(fos_handle_t h, int64_t i_start, int64_t n_iters, int64_t checki, double eps,
                      int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec);
/* fos_finish for every problem: guess B x fos_iterate_length, record B x FOS_REC_LEN, n_rec / status B. */
int32_t fos_finish_batch(fos_handle_t h, double *guess, double *record, int64_t *n_rec, int32_t *status);
/* solve!(model) for every problem in ONE kernel launch: begin, iterations 1..max_iters, getsol, final check. */
int32_t fos_solve_batch(fos_handle_t h, int64_t max_iters, int64_t checki, double eps, double *guess,
                        int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec);

/* ====================================================================================== */
/* unit-level entry points (drive the restated unit tests of the reference)  */
/* ====================================================================================== */
/* y = A x (transpose = 0, x: n, y: m) or y = A' x (transpose = 1, x: m, y: n). */
int32_t fos_a_mul(fos_handle_t h, const double *x, double *y, int32_t transpose);
/* Y = Q B / Q' B with Q = [0 A' c; -A 0 b; -c' -b' 0] (HSDEAffine.jl:41-65); length m+n+1. */
int32_t fos_q_mul(fos_handle_t h, const double *B, double *Y, int32_t transpose);
/* y = [I Op'; Op -I] x (affinepluslinear.jl:37-49); length = fos_iterate_length. */
int32_t fos_kkt_mul(fos_handle_t h, const double *x, double *y);
/* y = prox of S1 at x (affinepluslinear.jl:83-126), including every side effect (S1.i,
 * warm start, cgiter). */
int32_t fos_affine_prox(fos_handle_t h, const double *x, double *y);
/* y = HSDEMatrix.prox!(x) (problemforms/HSDE/HSDEAffine.jl:105-126): CG on [I Q'; Q -I] y = x with
 * the fixed tolerance 2l*eps, then v <- Q u.  Tested by the reference (test/HSDEAffine.jl:71-81)
 * but not used by its solver path; runs on a fresh CG state and leaves S1 untouched. */
int32_t fos_hsdematrix_prox(fos_handle_t h, const double *x, double *y);
/* y = prox of S2 at x (DualConeProduct cones.jl:122-142 / ConeProduct cones.jl:89-94). */
int32_t fos_cone_prox(fos_handle_t h, const double *x, double *y);
/* conjugategradient!(x, A, b, r, p, Ap; tol, max_iters) (conjugategradients.jl:31-55) on a
 * dense symmetric n x n matrix (row-major, host); x is the warm start and the result.
 * tol < 0 selects the reference default n*eps().  Stand-alone: needs no loaded problem. */
int32_t fos_cg_dense(fos_handle_t h, int64_t n, const double *A, const double *b, double *x, double tol,
                     int64_t max_iters, int64_t *iters);
/* Stand-alone projection onto one cone (dual != 0 -> proxDual!, cones.jl:80-102). */
int32_t fos_prox_cone(fos_handle_t h, int32_t cone_type, int32_t dual, const double *x, double *y, int64_t len);

/* Host-only: the work partition of the fused mat-vec for an m_local x n matrix over `ctas`
 * persistent CTAs (needs no GPU; unit-tested on CPU).  dims_out = {G, RT, NB, nslots, kc_last};
 * unit_begin has G+1 entries, slot_base NB+1, first_cta NB (pass capacities; -1 on overflow). */
int32_t fos_k1_plan(int64_t m_local, int64_t n, int32_t ctas, int32_t *dims_out, int32_t *unit_begin,
                    int64_t unit_begin_cap, int32_t *slot_base, int32_t *first_cta, int64_t band_cap);

/* Host-only: the kernel geometry batch mode picks for problems of shape m x n (needs no GPU; unit-tested on
 * CPU).  out = {lda, tiles of 8 rows, ring stages, consumer warps, column pairs per consumer thread, CTAs per
 * SM, dynamic shared memory bytes, doubles of A per problem}.  FOS_ERR_UNSUPPORTED for shapes outside batch mode. */
int32_t fos_batch_plan(int64_t m, int64_t n, int64_t *out /* 8 */);

/* Host-only: the decision of the hybrid row storage ("hybrid_rows") from the non-zeros per row (needs no GPU;
 * unit-tested on CPU).  out = {1 if the hybrid layout is used, first row of the dense block (a multiple of 16),
 * rows of the dense block, rows kept as CSR + CSC, their non-zeros}. */
int32_t fos_hybrid_plan(int64_t m, int64_t n, const int32_t *row_nnz, int64_t *out /* 5 */);

/* ====================================================================================== */
/* measurement helpers (bench.py)                                                         */
/* ====================================================================================== */
/* The CUDA stream (cudaStream_t as an integer) all work of this handle is enqueued on, so that a
 * caller can bracket calls with its own CUDA events (torch.cuda.ExternalStream). */
int32_t fos_get_stream(fos_handle_t h, uint64_t *stream_out);
/* Times `reps` launches of the fused dual mat-vec (nvec = 1 or 2 right-hand sides per
 * direction) on the loaded matrix with CUDA events on the library's stream; returns the
 * average milliseconds per launch and the algorithmic bytes one launch streams. */
/* With the option "tail_trace" = 1: summed SM cycles per phase of the fused CG-tail kernel, for its first, middle
 * and last block: out[16*b + k], k = 0 fold of the local partials, 1 block sync + system fence, 2 flags published
 * and peers seen (0..2 only with the peer-memory exchange), 3 gather + KKT epilogue, 4 first grid all-reduce,
 * 5 x / r update, 6 second grid all-reduce, 7 direction update; out[16*b + 15] = launches counted. */
int32_t fos_get_tail_trace(fos_handle_t h, double *out /* 48 */);
int32_t fos_time_matvec(fos_handle_t h, int32_t nvec, int32_t reps, double *ms_per_launch, double *bytes_per_launch);
/* Times `reps` projections of `ncones` packed symmetric matrices of order d (x: ncones * d(d+1)/2
 * doubles, host) onto the PSD cone (IndPSD(scaling=true), cones.jl:11) with CUDA events on the library's
 * stream; returns the average milliseconds per projection call (all cones in one launch), the number
 * of Jacobi sweeps of the last call (cone 0; 0 for the small-cone kernel) and, when y != NULL, the
 * projections.  Cold start every time (no warm start from the previous call's eigenvectors).  Stand-alone:
 * needs no loaded problem. */
int32_t fos_time_psd(fos_handle_t h, int64_t d, int64_t ncones, const double *x, double *y, int32_t reps,
                     double *ms_per_call, int32_t *sweeps);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FOS_B200_H */
