// common.cuh -- shared declarations of the fos_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fos_b200.h"

namespace fos {

// ---------------------------------------------------------------------------------------
// errors: C++ exceptions inside the library, converted to int32 codes at the C boundary
// ---------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &what) : std::runtime_error(what), code(c) {}
};

#define FOS_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            throw ::fos::Error(FOS_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + \
                                                 " (" __FILE__ ":" + std::to_string(__LINE__) + ")");    \
        }                                                                                                \
    } while (0)

// A synchronous cudaMemcpy from pageable host memory returns once the data sit in the driver's staging buffer;
// the DMA itself runs on the legacy default stream, which does not order against the handle's non-blocking
// stream.  Every synchronous host-to-device copy (and cudaMemset) is therefore followed by this wait.
#define FOS_SYNC_LEGACY() FOS_CUDA(cudaStreamSynchronize(cudaStreamLegacy))

#define FOS_REQUIRE(cond, msg)                                                   \
    do {                                                                         \
        if (!(cond)) throw ::fos::Error(FOS_ERR_INVALID, std::string(msg));      \
    } while (0)

// ---------------------------------------------------------------------------------------
// sizes
// ---------------------------------------------------------------------------------------
constexpr int64_t PAD = 16;  // every segment of a device vector is padded to 16 doubles (128 B)
inline int64_t ru(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

constexpr int MAX_PARTIALS = 1024;  // upper bound on the grid of any reduction kernel
constexpr int RED_SLOTS = 8;        // quantities reduced at once by one kernel

// ---------------------------------------------------------------------------------------
// device control block: every scalar of the CG recurrence and of the algorithms lives on
// the device, so a whole batch of CG iterations is enqueued without host round trips.
// ---------------------------------------------------------------------------------------
struct Ctrl {
    // conjugate gradients (utilities/conjugategradients.jl:31-55)
    double rn;      // r.r
    double alpha;   // rn / <Ap,p>
    double beta;    // rn / rnold
    double rnorm;   // ||r|| of the last update
    double tol;     // absolute tolerance of this solve
    int32_t iter;   // loop-body count (what conjugategradient! returns)
    int32_t max_iters;
    int32_t done;   // 1 -> every remaining CG kernel of the batch returns immediately
    int32_t warn_maxit;
    // GAPA (solvers/gapa.jl)
    double alpha12;
    // status check (HSDEStatus.jl / FeasibilityStatus.jl)
    int32_t status;
    int32_t nrec;       // records written so far in this run
    double feas_err;
    // GAPP line search
    double ls_normbest;
    double ls_alphabest;
    // scratch for scalar hand-off between kernels
    double scal[8];
    // device-side bookkeeping of the graph path (Handle::run_graph): no host round trip between iterations
    int64_t s1_calls_dev;   // S1.i (affinepluslinear.jl:66): indexes the tolerance table, advanced by k_iter_begin
    int64_t cur_i;          // iteration index (solverwrapper.jl:24), advanced by k_iter_begin
    int64_t total_cg_dev;   // CG iterations executed since the run started
    double fista_t;         // FISTA's t (fista.jl:14) while the graph path runs; the host copy is synchronised around it
    double fista_coef;      // (t_old - 1) / t of the current iteration (fista.jl:45-46), set by k_iter_begin
    uint32_t p2p_error;  // host copy only: peer-exchange timeout flag fetched by sync_ctrl
    uint32_t pad_;
};

// deterministic two-level reduction workspace
struct RedBuf {
    double *partials;       // [RED_SLOTS][MAX_PARTIALS]
    unsigned int *counter;  // last-block ticket, self-resetting
};

// ---------------------------------------------------------------------------------------
// Result view of one dual mat-vec: per-band row partials and per-slot column partials.
//   (A X)[v][i]   = sum_{b < nb}  rowpart[b*rp_sb + v*rp_sv + i]
//   (A' W)[v][j]  = sum_{s in [slot_base[band], slot_base[band+1])} colpart[s*cp_ss + v*cp_sv + (j & bw_mask)]
//                   with band = j >> bw_shift
// A complete (already reduced) result is the special case nb = 1, one band, one slot.
// ---------------------------------------------------------------------------------------
struct MVView {
    const double *rowpart;
    int32_t nb;
    int64_t rp_sb, rp_sv;
    const double *colpart;
    const int32_t *slot_base;
    int32_t bw_shift;
    int64_t bw_mask;
    int64_t cp_ss, cp_sv;
};

#ifdef __CUDACC__
// The partials of one entry are added in index order; loads are issued eight at a time so that their
// L2 latencies overlap (a plain running sum serialises ~20 dependent round trips at config 2).
__device__ __forceinline__ double mv_ax(const MVView &V, int v, int64_t i)
{
    double s = 0.0;
    const double *p = V.rowpart + v * V.rp_sv + i;
    for (int b0 = 0; b0 < V.nb; b0 += 8) {
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = (b0 + k < V.nb) ? p[(int64_t)(b0 + k) * V.rp_sb] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += t[k];
    }
    return s;
}
__device__ __forceinline__ double mv_atw(const MVView &V, int v, int64_t j)
{
    int64_t band = j >> V.bw_shift;
    int64_t jj = j & V.bw_mask;
    int s0 = V.slot_base[band], s1 = V.slot_base[band + 1];
    double s = 0.0;
    const double *p = V.colpart + v * V.cp_sv + jj;
    for (int t0 = s0; t0 < s1; t0 += 8) {
        double t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) t[k] = (t0 + k < s1) ? p[(int64_t)(t0 + k) * V.cp_ss] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += t[k];
    }
    return s;
}

// Both right-hand sides at once: the loads of v = 0 and v = 1 travel together (half the dependent round trips
// of two mv_ax / mv_atw calls); each sum is still added in index order -> bitwise equal to the single versions.
__device__ __forceinline__ void mv_ax2(const MVView &V, int64_t i, double &s0, double &s1)
{
    s0 = 0.0;
    s1 = 0.0;
    const double *p0 = V.rowpart + i, *p1 = V.rowpart + V.rp_sv + i;
    for (int b0 = 0; b0 < V.nb; b0 += 8) {
        double t0[8], t1[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const bool in = b0 + k < V.nb;
            t0[k] = in ? p0[(int64_t)(b0 + k) * V.rp_sb] : 0.0;
            t1[k] = in ? p1[(int64_t)(b0 + k) * V.rp_sb] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            s0 += t0[k];
            s1 += t1[k];
        }
    }
}
__device__ __forceinline__ void mv_atw2(const MVView &V, int64_t j, double &s0, double &s1)
{
    const int64_t band = j >> V.bw_shift;
    const int64_t jj = j & V.bw_mask;
    const int a = V.slot_base[band], b = V.slot_base[band + 1];
    s0 = 0.0;
    s1 = 0.0;
    const double *p0 = V.colpart + jj, *p1 = V.colpart + V.cp_sv + jj;
    for (int t0i = a; t0i < b; t0i += 8) {
        double t0[8], t1[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const bool in = t0i + k < b;
            t0[k] = in ? p0[(int64_t)(t0i + k) * V.cp_ss] : 0.0;
            t1[k] = in ? p1[(int64_t)(t0i + k) * V.cp_ss] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            s0 += t0[k];
            s1 += t1[k];
        }
    }
}

// no-contraction arithmetic: the reference's broadcasts round every operation separately
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level deterministic reduction of NQ quantities followed by the "last block finishes"
// pattern.  Returns true in every thread of the last block to arrive; in that block tot[q]
// (thread 0 only) holds the grid-wide sums.  The summation tree has a fixed shape for a given
// grid, so results are bitwise reproducible run to run (no floating-point atomics).
template <int NQ, int BLOCK>
__device__ __forceinline__ bool grid_reduce(const double (&val)[NQ], double (&tot)[NQ], const RedBuf &rb)
{
    __shared__ double s_w[NQ][BLOCK / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        double v = warp_sum(val[q]);
        if (lane == 0) s_w[q][w] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            double s = 0.0;
            for (int k = 0; k < BLOCK / 32; k++) s += s_w[q][k];
            rb.partials[q * MAX_PARTIALS + blockIdx.x] = s;
        }
        __threadfence();
        unsigned int t = atomicAdd(rb.counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const volatile double *p = rb.partials + q * MAX_PARTIALS;
        double acc = 0.0;
        for (unsigned int k = threadIdx.x; k < gridDim.x; k += BLOCK) acc += p[k];
        acc = warp_sum(acc);
        if (lane == 0) s_w[q][w] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            double s = 0.0;
            for (int k = 0; k < BLOCK / 32; k++) s += s_w[q][k];
            tot[q] = s;
        }
        *rb.counter = 0u;  // self-reset for the next kernel on the stream
    }
    return true;
}
#endif  // __CUDACC__

}  // namespace fos
