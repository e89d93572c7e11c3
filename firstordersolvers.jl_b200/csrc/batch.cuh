// batch.cuh -- batch mode (config 5): B independent conic problems of identical shape (same m, n and
// cone layout, different A, b, c), each solved start to finish by ONE persistent CTA.
//
// Replaces, per problem, exactly what the single-problem path replaces (solverwrapper.jl:20-41 and
// everything below it), but with no host round trips at all: the outer loop, the CG loop with its
// stop test, the cone projections and the residual checks run inside one kernel launch.  CTAs pull
// problems from a work counter, so problems that stop early (or need fewer CG iterations) do not hold
// the others back -- there are no per-problem masks.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace fos {

constexpr int BT_TR = 8;            // rows per A tile (NV * TR = 16 row sums per tile).  4-row tiles (four 17 KB stages at
                                    // config 5) were measured 12 % SLOWER: the per-tile barrier + reduce dominates (profiles/r2_c5_notes.md)
constexpr int BT_GROUP = 4;         // tiles between two meetings of the consumer warps (power of two)
constexpr int BT_MAX_STAGES = 4;
constexpr int BT_MAX_CW = 15;       // consumer warps
constexpr int BT_MAX_KP = 4;        // column pairs per consumer thread
constexpr int BT_MAX_SOC = 16;      // SOC cones per problem (2 per SOC of K1/K2: primal and dual image)
constexpr int BT_MAX_LDA = 1280;

// per-problem vectors in the workspace, each NP doubles (padded layout of solver.cuh)
enum { BV_X = 0, BV_SOL, BV_R, BV_P, BV_AP, BV_TMP1, BV_TMP2, BV_PROJ, BV_FY, BV_FXOLD, BV_DP, BV_DQ, BV_DY, BV_W1,
       BV_W2, BV_COUNT };

struct BatchCtl {  // persistent per-problem state (what GAPData/.../AffinePlusLinear/HSDEStatus hold)
    int64_t s1_calls;   // affinepluslinear.jl:66, starts at 1
    int32_t firstrun;   // :101-104
    int32_t cgiter;     // :67 / :121
    int32_t status;     // HSDEStatus
    int32_t checked;
    int32_t warn_maxit;
    int32_t nrec;       // records written by the current call
    int64_t last_i;     // last executed iteration (status.i)
    int64_t iters_done; // iterations executed by the current call
    int64_t total_cg, total_passes;
    double alpha12;     // gapa.jl:18
    double fista_t;     // fista.jl:14
};

struct BatchArgs {
    Lay L;
    int64_t lda;        // leading dimension of every A (doubles, multiple of 16)
    int32_t ntiles;     // tiles of BT_TR rows per problem
    int32_t S, CW, KP;  // ring stages, consumer warps, column pairs per consumer thread
    const double *A;    // [B][ntiles*BT_TR][lda]: the DENSE rows of every problem, gathered into tiles
    int64_t a_stride;
    // Hybrid storage: rows with few non-zeros (<= n/8) are kept out of the dense tiles and stored as CSR (for
    // A x) + CSC (for A' w); all-zero rows are stored nowhere.  Config 5's conic NNLS form has 256 dense rows
    // and 513 rows with a single entry: 1.08 MB of tiles instead of 3.2 MB per problem and pass.
    const int32_t *drow;      // [B][ntiles*BT_TR] original row of every tile row (mr - 16 = dummy slot for padding)
    const int32_t *sp_count;  // [B][2] {sparse rows, columns with sparse entries}
    const int32_t *srow_ptr, *srow_id, *scol;  // [B][NS+1], [B][NS], [B][NZ]
    const double *sval;                        // [B][NZ]
    const int32_t *ccol_ptr, *ccol_id, *crow;  // [B][NCS+1], [B][NCS], [B][NZ]
    const double *cval;                        // [B][NZ]
    int32_t NS, NCS, NZ;                       // capacities = strides of the arrays above
    int32_t mr;                                // row slots of the shared-memory A X / W arrays (m_pad + 16)
    const double *b, *c;    // [B][m_pad], [B][n_pad]
    const double *nb, *ncn; // [B] ||b||, ||c||
    double *vec;            // [B][BV_COUNT][NP]
    BatchCtl *ctl;          // [B]
    double *recs;           // [B][rec_cap][FOS_REC_LEN]
    int32_t rec_cap;
    const double *tol_table;  // tol_table[i-1] = max(0.2^sqrt(i), l*eps), i = 1..tol_n; beyond: tol_floor
    int32_t tol_n;
    double tol_floor;
    const uint8_t *ops;       // [NP] cone op codes (shared by all problems)
    const int32_t *cone_of;   // [NP]
    const SocCone *soc;       // [nsoc]
    int32_t nsoc;
    unsigned int *work_counter;
    int32_t B;
    int32_t alg;
    double alpha, alpha1, alpha2, betaA;
    int64_t i_start, n_iters, checki;
    double eps;
    int32_t do_run, do_finish;
};

struct BatchGeom {
    int64_t lda, a_stride;
    int32_t ntiles, S, CW, KP, ctas_per_sm;
    size_t smem_bytes;
};
BatchGeom batch_geometry(int64_t m, int64_t n);  // host only; throws FOS_ERR_UNSUPPORTED for shapes outside batch mode

}  // namespace fos
