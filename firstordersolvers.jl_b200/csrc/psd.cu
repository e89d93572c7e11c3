// psd.cu -- K5: projection onto the PSD cone, IndPSD(scaling=true) of ProximalOperators.jl as
// mapped at cones.jl:11 (dual through Moreau, cones.jl:80-85).
//
// Input/output: packed lower triangle, column-major, off-diagonals carrying the sqrt(2) svec
// scaling (MathProgBase).  Algorithm: diag *= sqrt2 -> symmetric eigendecomposition -> clamp
// the eigenvalues at 0 -> V max(L,0) V' -> repack, diag /= sqrt2.
//
// Eigendecomposition = parallel two-sided (symmetric) Jacobi with round-robin pair ordering:
// each of the D-1 steps of a sweep applies D/2 disjoint plane rotations at once
// (S <- J' S J, V <- V J).  One CTA per cone; S and V live in shared memory (2*d*d*8 bytes, d <= 112).
// Larger cones take the cooperative multi-CTA path of psd_large.cu.
#include <algorithm>

#include "solver.cuh"

namespace fos {

constexpr int PSD_THREADS = 512;
constexpr int PSD_MAX_SWEEPS = 40;

__device__ __forceinline__ void rr_pair(int s, int k, int D, int &a, int &b)
{
    // round-robin tournament: player D-1 is fixed, the others rotate
    const int M = D - 1;
    if (k == 0) {
        a = D - 1;
        b = s;
    } else {
        a = (s + k) % M;
        b = (s - k + M) % M;
    }
    if (a > b) {
        const int t = a;
        a = b;
        b = t;
    }
}

__global__ void __launch_bounds__(PSD_THREADS)
k5_psd_jacobi(const PsdCone *__restrict__ cones, const double *__restrict__ in, double *__restrict__ proj,
              double *__restrict__ work, int64_t work_stride, int use_smem)
{
    extern __shared__ double psd_smem[];
    __shared__ int s_rot;
    __shared__ double s_red[PSD_THREADS / 32];
    __shared__ double s_thr;
    const PsdCone C = cones[blockIdx.x];
    const int d = C.d;
    const int D = (d + 1) & ~1;  // even number of players; index d (if any) is a dummy
    double *S = use_smem ? psd_smem : work + (size_t)blockIdx.x * work_stride;
    double *V = S + (size_t)d * d;
    double *cs = V + (size_t)d * d;  // [D/2][2]
    const double sq2 = 1.4142135623730951;
    const double sgn = C.dual ? -1.0 : 1.0;  // dual: project -x, then add x

    // ---- unpack ----
    for (int64_t idx = threadIdx.x; idx < (int64_t)d * d; idx += PSD_THREADS) {
        const int i = (int)(idx / d), j = (int)(idx % d);
        const int lo = i > j ? i : j, hi = i > j ? j : i;  // lower triangle element (lo, hi), column hi
        const int64_t k = (int64_t)hi * d - (int64_t)hi * (hi - 1) / 2 + (lo - hi);
        double v = sgn * in[C.off + k];
        if (i == j) v *= sq2;
        S[idx] = v;
        V[idx] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    // rotation threshold: |a_pq| > 1e-17 * ||S||_F (entries below it cannot move an eigenvalue by more
    // than rounding); annihilated entries are set to exactly 0, so the iteration terminates.
    {
        double acc = 0.0;
        for (int64_t idx = threadIdx.x; idx < (int64_t)d * d; idx += PSD_THREADS) acc = fma(S[idx], S[idx], acc);
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < PSD_THREADS / 32; k++) t += s_red[k];
            s_thr = 1e-17 * sqrt(t);
        }
        __syncthreads();
    }
    const double thr = s_thr;

    // ---- Jacobi sweeps ----
    const int npairs = D / 2;
    for (int sweep = 0; sweep < PSD_MAX_SWEEPS; sweep++) {
        if (threadIdx.x == 0) s_rot = 0;
        __syncthreads();
        for (int step = 0; step < D - 1; step++) {
            // phase 1: rotation parameters from the current S
            for (int k = threadIdx.x; k < npairs; k += PSD_THREADS) {
                int p, q;
                rr_pair(step, k, D, p, q);
                double c = 1.0, s = 0.0;
                if (q < d) {
                    const double apq = S[(size_t)p * d + q];
                    const double app = S[(size_t)p * d + p], aqq = S[(size_t)q * d + q];
                    if (fabs(apq) > thr) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                        s_rot = 1;
                    }
                }
                cs[2 * k] = c;
                cs[2 * k + 1] = s;
            }
            __syncthreads();
            // phase 2: columns  S <- S J ,  V <- V J
            for (int idx = threadIdx.x; idx < npairs * d; idx += PSD_THREADS) {
                const int k = idx / d, row = idx - k * d;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                if (s == 0.0) continue;
                int p, q;
                rr_pair(step, k, D, p, q);
                const size_t ip = (size_t)row * d + p, iq = (size_t)row * d + q;
                const double sp = S[ip], sq = S[iq];
                S[ip] = c * sp - s * sq;
                S[iq] = s * sp + c * sq;
                const double vp = V[ip], vq = V[iq];
                V[ip] = c * vp - s * vq;
                V[iq] = s * vp + c * vq;
            }
            __syncthreads();
            // phase 3: rows  S <- J' S
            for (int idx = threadIdx.x; idx < npairs * d; idx += PSD_THREADS) {
                const int k = idx / d, col = idx - k * d;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                if (s == 0.0) continue;
                int p, q;
                rr_pair(step, k, D, p, q);
                const size_t ip = (size_t)p * d + col, iq = (size_t)q * d + col;
                const double sp = S[ip], sq = S[iq];
                S[ip] = (col == q) ? 0.0 : c * sp - s * sq;  // a_pq := 0 exactly
                S[iq] = (col == p) ? 0.0 : s * sp + c * sq;  // a_qp := 0 exactly
            }
            __syncthreads();
        }
        if (s_rot == 0) break;
        __syncthreads();
    }

    // ---- eigenvalues -> clamp; scale the columns of V by sqrt(lambda+) so P = W W' ----
    for (int idx = threadIdx.x; idx < d * d; idx += PSD_THREADS) {
        const int e = idx % d;
        const double lam = S[(size_t)e * d + e];
        // keep lambda on the diagonal until every thread has read it: write scaled V to itself only
        V[idx] = lam > 0.0 ? V[idx] * sqrt(lam) : 0.0;
    }
    __syncthreads();
    // ---- P = W W' (lower triangle), repack, Moreau for the dual ----
    const int64_t plen = (int64_t)d * (d + 1) / 2;
    for (int64_t k = threadIdx.x; k < plen; k += PSD_THREADS) {
        // invert k -> (col j, row i) of the packed lower triangle
        int j = (int)floor(((2.0 * d + 1.0) - sqrt((2.0 * d + 1.0) * (2.0 * d + 1.0) - 8.0 * (double)k)) / 2.0);
        while ((int64_t)j * d - (int64_t)j * (j - 1) / 2 > k) j--;
        while ((int64_t)(j + 1) * d - (int64_t)(j + 1) * j / 2 <= k) j++;
        const int i = j + (int)(k - ((int64_t)j * d - (int64_t)j * (j - 1) / 2));
        double acc = 0.0;
        const double *wi = V + (size_t)i * d, *wj = V + (size_t)j * d;
        for (int e = 0; e < d; e++) acc = fma(wi[e], wj[e], acc);
        if (i == j) acc /= sq2;
        const double x = in[C.off + k];
        proj[C.off + k] = C.dual ? __dadd_rn(x, acc) : acc;
    }
}

// ---------------------------------------------------------------------------------------
// Tiny cones (d <= PSD_WARP_MAX_D = 32; the 2x2 / 3x3 blocks of LMI models, hundreds of them): ONE WARP per cone,
// PSD_WPB cones per CTA.  Same two-sided Jacobi and round-robin order as the kernel above, but S and V of a cone sit
// in the warp's slice of shared memory, the three phases of a step are separated by __syncwarp only (no CTA barrier,
// no idle 480 threads), a lane owns a whole row (column) in the update phases, the pair schedule is tabulated once
// (no integer division in the sweeps) and the rotation needs one sqrt, one divide and one rsqrt.
// 1024 cones of order 16: 0.68 ms with one 512-thread CTA per cone (round 1) against 0.46 ms for cuSOLVER's batched
// syevj; this kernel: profiles/r2_psd_probe.jsonl.
// ---------------------------------------------------------------------------------------
constexpr int PSD_WPB = 8;

// shared memory of one warp: S and V with an ODD leading dimension (row stride d or d+1: the column phase reads
// S[lane][p] -- an even stride would put all lanes on one bank), the (c, s) pairs of a step and the round-robin
// schedule (p, q) of every (step, pair), built once so that the sweeps contain no integer division
__host__ __device__ inline int psd_warp_ld(int d) { return d | 1; }
__host__ __device__ inline size_t psd_warp_doubles(int dmax)
{
    const int D = (dmax + 1) & ~1, ld = psd_warp_ld(dmax);
    const size_t tab_bytes = (size_t)(D - 1) * (D / 2) * 2;
    return (size_t)2 * dmax * ld + (size_t)D + (tab_bytes + 7) / 8;
}

__global__ void __launch_bounds__(PSD_WPB * 32)
k5_psd_jacobi_warp(const PsdCone *__restrict__ cones, int ncones, const double *__restrict__ in,
                   double *__restrict__ proj, int dmax)
{
    extern __shared__ double psd_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = blockIdx.x * PSD_WPB + warp;
    if (ci >= ncones) return;  // whole warps leave; nothing below synchronises across warps
    const PsdCone C = cones[ci];
    const int d = C.d;
    const int D = (d + 1) & ~1;
    const int ld = psd_warp_ld(d);
    double *S = psd_smem + (size_t)warp * psd_warp_doubles(dmax);
    double *V = S + (size_t)dmax * psd_warp_ld(dmax);
    double *cs = V + (size_t)dmax * psd_warp_ld(dmax);
    unsigned char *tab = reinterpret_cast<unsigned char *>(cs + ((dmax + 1) & ~1));
    const double sq2 = 1.4142135623730951;
    const double sgn = C.dual ? -1.0 : 1.0;
    const int npairs = D / 2;
    for (int idx = lane; idx < (D - 1) * npairs; idx += 32) {
        int p, q;
        rr_pair(idx / npairs, idx % npairs, D, p, q);
        tab[2 * idx] = (unsigned char)p;
        tab[2 * idx + 1] = (unsigned char)q;
    }
    double acc = 0.0;
    for (int idx = lane; idx < d * d; idx += 32) {
        const int i = idx / d, j = idx - i * d;
        const int lo = i > j ? i : j, hi = i > j ? j : i;
        const int64_t k = (int64_t)hi * d - (int64_t)hi * (hi - 1) / 2 + (lo - hi);
        double v = sgn * in[C.off + k];
        if (i == j) v *= sq2;
        S[i * ld + j] = v;
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
        acc = fma(v, v, acc);
    }
    const double thr = 1e-17 * sqrt(warp_sum(acc));
    __syncwarp();
    // update phases: lane -> (row, pair group).  With d <= 16 a row is shared by 32 / P lanes (P = d rounded up to a
    // power of two), each taking every G-th pair of the step: the pairs are disjoint, so the groups touch different columns
    int P = 1;
    while (P < d) P <<= 1;
    const int G = 32 / P, row = lane & (P - 1), grp = lane / P;
    for (int sweep = 0; sweep < PSD_MAX_SWEEPS; sweep++) {
        bool rotated = false;
        for (int step = 0; step < D - 1; step++) {
            const unsigned char *tp = tab + 2 * step * npairs;
            // phase 1: one lane per pair.  t = tan of the rotation angle from the stable two-term form
            // t = 2 a_pq / (delta + sign(delta) sqrt(delta^2 + 4 a_pq^2)): one sqrt, one divide, one rsqrt
            bool mine = false;
            if (lane < npairs) {
                const int p = tp[2 * lane], q = tp[2 * lane + 1];
                double c = 1.0, sn = 0.0;
                if (q < d) {
                    const double apq = S[p * ld + q];
                    if (fabs(apq) > thr) {
                        const double delta = S[q * ld + q] - S[p * ld + p];
                        const double rt = sqrt(fma(delta, delta, 4.0 * apq * apq));
                        const double t = (2.0 * apq) / (delta >= 0.0 ? delta + rt : delta - rt);
                        c = rsqrt(fma(t, t, 1.0));
                        sn = t * c;
                        mine = true;
                    }
                }
                cs[2 * lane] = c;
                cs[2 * lane + 1] = sn;
            }
            rotated = rotated || __any_sync(0xffffffffu, mine);
            __syncwarp();
            // phase 2: (row, group) lanes.  S <- S J, V <- V J  (the pairs of a step are disjoint: independent updates)
            if (row < d) {
                double *Sr = S + row * ld, *Vr = V + row * ld;
#pragma unroll 4
                for (int k = grp; k < npairs; k += G) {
                    const double c = cs[2 * k], sn = cs[2 * k + 1];
                    if (sn == 0.0) continue;
                    const int p = tp[2 * k], q = tp[2 * k + 1];
                    const double sp = Sr[p], sq = Sr[q];
                    Sr[p] = c * sp - sn * sq;
                    Sr[q] = sn * sp + c * sq;
                    const double vp = Vr[p], vq = Vr[q];
                    Vr[p] = c * vp - sn * vq;
                    Vr[q] = sn * vp + c * vq;
                }
            }
            __syncwarp();
            // phase 3: (column, group) lanes.  S <- J' S, the annihilated entries set to exactly 0
            if (row < d) {
#pragma unroll 4
                for (int k = grp; k < npairs; k += G) {
                    const double c = cs[2 * k], sn = cs[2 * k + 1];
                    if (sn == 0.0) continue;
                    const int p = tp[2 * k], q = tp[2 * k + 1];
                    const double sp = S[p * ld + row], sq = S[q * ld + row];
                    S[p * ld + row] = (row == q) ? 0.0 : c * sp - sn * sq;
                    S[q * ld + row] = (row == p) ? 0.0 : sn * sp + c * sq;
                }
            }
            __syncwarp();
        }
        if (!rotated) break;
    }
    // W = V sqrt(max(lambda, 0)) column-wise, then P = W W'
    if (lane < d) {
        for (int e = 0; e < d; e++) {
            const double lam = S[e * ld + e];
            V[lane * ld + e] = lam > 0.0 ? V[lane * ld + e] * sqrt(lam) : 0.0;
        }
    }
    __syncwarp();
    const int plen = d * (d + 1) / 2;
    for (int k = lane; k < plen; k += 32) {
        int j = (int)floor(((2.0 * d + 1.0) - sqrt((2.0 * d + 1.0) * (2.0 * d + 1.0) - 8.0 * (double)k)) / 2.0);
        while (j * d - j * (j - 1) / 2 > k) j--;
        while ((j + 1) * d - (j + 1) * j / 2 <= k) j++;
        const int i = j + (k - (j * d - j * (j - 1) / 2));
        double a2 = 0.0;
        const double *wi = V + i * ld, *wj = V + j * ld;
        for (int e = 0; e < d; e++) a2 = fma(wi[e], wj[e], a2);
        if (i == j) a2 /= sq2;
        const double x = in[C.off + k];
        proj[C.off + k] = C.dual ? __dadd_rn(x, a2) : a2;
    }
}

void psd_project(Handle *h, ConeSet &K, const double *in, double *projbuf)
{
    const int nc = (int)K.psd.size();
    if (nc == 0) return;
    const int ns = K.psd_nsmall;
    if (ns > 0) {
        const int dm = K.psd_small_max_d;
        const size_t need_w = psd_warp_doubles(dm) * sizeof(double) * PSD_WPB;
        FOS_CUDA(cudaFuncSetAttribute(k5_psd_jacobi_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need_w));
        FOS_LAUNCH(h, k5_psd_jacobi_warp, (ns + PSD_WPB - 1) / PSD_WPB, PSD_WPB * 32, need_w, K.d_psd.p, ns, in, projbuf, dm);
    }
    if (nc > ns) {
        const int d = K.psd_max_d;
        const int D = (d + 1) & ~1;
        const size_t need = ((size_t)2 * d * d + (size_t)D) * sizeof(double);
        FOS_CUDA(cudaFuncSetAttribute(k5_psd_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        FOS_LAUNCH(h, k5_psd_jacobi, nc - ns, PSD_THREADS, need, K.d_psd.p + ns, in, projbuf, (double *)nullptr, (int64_t)0, 1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(FOS_ERR_CUDA, std::string("PSD projection launch failed: ") + cudaGetErrorString(e));
}

}  // namespace fos
