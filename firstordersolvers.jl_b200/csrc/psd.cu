// psd.cu -- K5: projection onto the PSD cone, IndPSD(scaling=true) of ProximalOperators.jl as
// mapped at cones.jl:11 (dual through Moreau, cones.jl:80-85).
//
// Input/output: packed lower triangle, column-major, off-diagonals carrying the sqrt(2) svec
// scaling (MathProgBase).  Algorithm: diag *= sqrt2 -> symmetric eigendecomposition -> clamp
// the eigenvalues at 0 -> V max(L,0) V' -> repack, diag /= sqrt2.
//
// Eigendecomposition = parallel two-sided (symmetric) Jacobi with round-robin pair ordering:
// each of the D-1 steps of a sweep applies D/2 disjoint plane rotations at once
// (S <- J' S J, V <- V J).  S and V of a cone live in shared memory (d <= 112); 32, 128 or 512 threads work on a cone
// depending on its order (k5_psd_jacobi_group).  Larger cones take the cooperative multi-CTA path of psd_large.cu.
#include <algorithm>

#include "solver.cuh"

namespace fos {

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

// ---------------------------------------------------------------------------------------
// k5_psd_jacobi_group<TPC>: TPC threads per cone, 256 / TPC (or one, TPC = 512) cones per CTA.
//   d <= 16      TPC = 32  : one warp per cone, eight cones per CTA, __syncwarp between the phases
//   d <= 48      TPC = 128 : two cones per CTA, one named barrier per cone
//   d <= 112     TPC = 512 : one cone per CTA
// S and V of a cone sit in the group's slice of shared memory with an ODD leading dimension (the column phase reads
// S[row][p]: an even stride would put all rows on one bank).  The round-robin schedule (p, q) of every (step, pair) is
// tabulated once, so the sweeps contain no integer division.  A step has three phases:
//   1. one thread per pair: rotation (c, s) from t = 2 a_pq / (delta + sign(delta) sqrt(delta^2 + 4 a_pq^2)),
//      c = rsqrt(1 + t^2), s = t c  -- one sqrt, one divide, one rsqrt;
//   2. thread -> (row, pair group): S <- S J, V <- V J (the pairs of a step are disjoint: independent updates);
//   3. thread -> (column, pair group): S <- J' S, the annihilated entries set to exactly 0.
// Round 1 ran one 512-thread CTA per cone with a flat element loop (two integer divisions and two modulo per element
// update): 1024 cones of order 16 took 0.68 ms against 0.46 ms for cuSOLVER's batched syevj; this kernel 0.19 ms.
// ---------------------------------------------------------------------------------------
__host__ __device__ inline int psd_group_ld(int d) { return d | 1; }
__host__ __device__ inline size_t psd_group_doubles(int dmax)
{
    const int D = (dmax + 1) & ~1, ld = psd_group_ld(dmax);
    const size_t tab_bytes = (size_t)(D - 1) * (D / 2) * 2;
    return (size_t)2 * dmax * ld + (size_t)D + (tab_bytes + 7) / 8 + 1;  // S, V, (c, s) pairs, schedule, rotation flag
}

template <int TPC>
__device__ __forceinline__ void psd_group_sync(int bar_id)
{
    if (TPC == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TPC) : "memory");
}

template <int TPC>
__global__ void __launch_bounds__(TPC >= 256 ? TPC : 256)
k5_psd_jacobi_group(const PsdCone *__restrict__ cones, int ncones, const double *__restrict__ in,
                    double *__restrict__ proj, int dmax)
{
    constexpr int CTA = TPC >= 256 ? TPC : 256, GPC = CTA / TPC;  // groups (cones) per CTA
    extern __shared__ double psd_smem[];
    const int grp_in_cta = threadIdx.x / TPC, t = threadIdx.x % TPC;
    const int bar_id = 1 + grp_in_cta;
    const int ci = blockIdx.x * GPC + grp_in_cta;
    if (ci >= ncones) return;  // whole groups leave; nothing below synchronises across groups
    const PsdCone C = cones[ci];
    const int d = C.d;
    const int D = (d + 1) & ~1;
    const int ld = psd_group_ld(d);
    double *S = psd_smem + (size_t)grp_in_cta * psd_group_doubles(dmax);
    double *V = S + (size_t)dmax * psd_group_ld(dmax);
    double *cs = V + (size_t)dmax * psd_group_ld(dmax);
    unsigned char *tab = reinterpret_cast<unsigned char *>(cs + ((dmax + 1) & ~1));
    volatile int *rot = reinterpret_cast<volatile int *>(S + psd_group_doubles(dmax) - 1);
    __shared__ double s_thr[8];
    __shared__ double s_acc[8][16];
    const double sq2 = 1.4142135623730951;
    const double sgn = C.dual ? -1.0 : 1.0;  // dual: project -x, then add x
    const int npairs = D / 2;
    for (int idx = t; idx < (D - 1) * npairs; idx += TPC) {
        int p, q;
        rr_pair(idx / npairs, idx % npairs, D, p, q);
        tab[2 * idx] = (unsigned char)p;
        tab[2 * idx + 1] = (unsigned char)q;
    }
    // ---- unpack; rotation threshold |a_pq| > 1e-17 ||S||_F (smaller entries cannot move an eigenvalue by more than
    // rounding; annihilated entries are set to exactly 0, so the iteration terminates) ----
    double acc = 0.0;
    for (int idx = t; idx < d * d; idx += TPC) {
        const int i = idx / d, j = idx - i * d;
        const int lo = i > j ? i : j, hi = i > j ? j : i;  // lower triangle element (lo, hi), column hi
        const int64_t k = (int64_t)hi * d - (int64_t)hi * (hi - 1) / 2 + (lo - hi);
        double v = sgn * in[C.off + k];
        if (i == j) v *= sq2;
        S[i * ld + j] = v;
        V[i * ld + j] = (i == j) ? 1.0 : 0.0;
        acc = fma(v, v, acc);
    }
    acc = warp_sum(acc);
    if (TPC == 32) {
        if (t == 0) s_thr[grp_in_cta] = 1e-17 * sqrt(acc);
    } else {
        if ((t & 31) == 0) s_acc[grp_in_cta][t >> 5] = acc;
        psd_group_sync<TPC>(bar_id);
        if (t == 0) {
            double tot = 0.0;
            for (int w = 0; w < TPC / 32; w++) tot += s_acc[grp_in_cta][w];
            s_thr[grp_in_cta] = 1e-17 * sqrt(tot);
        }
    }
    psd_group_sync<TPC>(bar_id);
    const double thr = s_thr[grp_in_cta];
    // update phases: thread -> (row, pair group); P = d rounded up to a power of two
    int P = 1;
    while (P < d) P <<= 1;
    const int G = TPC / P, row = t & (P - 1), grp = t / P;  // TPC >= P by the dispatch below
    for (int sweep = 0; sweep < PSD_MAX_SWEEPS; sweep++) {
        if (t == 0) *rot = 0;
        psd_group_sync<TPC>(bar_id);
        for (int step = 0; step < D - 1; step++) {
            const unsigned char *tp = tab + 2 * step * npairs;
            if (t < npairs) {
                const int p = tp[2 * t], q = tp[2 * t + 1];
                double c = 1.0, sn = 0.0;
                if (q < d) {
                    const double apq = S[p * ld + q];
                    if (fabs(apq) > thr) {
                        const double delta = S[q * ld + q] - S[p * ld + p];
                        const double rt = sqrt(fma(delta, delta, 4.0 * apq * apq));
                        const double tt = (2.0 * apq) / (delta >= 0.0 ? delta + rt : delta - rt);
                        c = rsqrt(fma(tt, tt, 1.0));
                        sn = tt * c;
                        *rot = 1;
                    }
                }
                cs[2 * t] = c;
                cs[2 * t + 1] = sn;
            }
            psd_group_sync<TPC>(bar_id);
            if (row < d && grp < G) {
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
            psd_group_sync<TPC>(bar_id);
            if (row < d && grp < G) {
#pragma unroll 4
                for (int k = grp; k < npairs; k += G) {
                    const double c = cs[2 * k], sn = cs[2 * k + 1];
                    if (sn == 0.0) continue;
                    const int p = tp[2 * k], q = tp[2 * k + 1];
                    const double sp = S[p * ld + row], sq = S[q * ld + row];
                    S[p * ld + row] = (row == q) ? 0.0 : c * sp - sn * sq;  // a_pq := 0 exactly
                    S[q * ld + row] = (row == p) ? 0.0 : sn * sp + c * sq;  // a_qp := 0 exactly
                }
            }
            psd_group_sync<TPC>(bar_id);
        }
        const int any = *rot;
        psd_group_sync<TPC>(bar_id);  // everybody has read the flag before thread 0 clears it again
        if (!any) break;
    }
    // ---- W = V sqrt(max(lambda, 0)) column-wise, then P = W W' (lower triangle), repack, Moreau for the dual ----
    for (int idx = t; idx < d * d; idx += TPC) {
        const int i = idx / d, e = idx - i * d;
        const double lam = S[e * ld + e];
        V[i * ld + e] = lam > 0.0 ? V[i * ld + e] * sqrt(lam) : 0.0;
    }
    psd_group_sync<TPC>(bar_id);
    const int plen = d * (d + 1) / 2;
    for (int k = t; k < plen; k += TPC) {
        // invert k -> (col j, row i) of the packed lower triangle
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

template <int TPC>
static void psd_launch_group(Handle *h, const PsdCone *cones, int nc, int dmax, const double *in, double *projbuf)
{
    if (nc <= 0) return;
    constexpr int CTA = TPC >= 256 ? TPC : 256, GPC = CTA / TPC;
    const size_t need = psd_group_doubles(dmax) * sizeof(double) * GPC;
    FOS_CUDA(cudaFuncSetAttribute(k5_psd_jacobi_group<TPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    FOS_LAUNCH(h, k5_psd_jacobi_group<TPC>, (nc + GPC - 1) / GPC, CTA, need, cones, nc, in, projbuf, dmax);
}

// cones are sorted by order at load (ConeSet::build): [0, n16) take a warp each, [n16, n48) 128 threads, the rest 512
void psd_project(Handle *h, ConeSet &K, const double *in, double *projbuf)
{
    const int nc = (int)K.psd.size();
    if (nc == 0) return;
    const int n1 = K.psd_n16, n2 = K.psd_n48;
    psd_launch_group<32>(h, K.d_psd.p, n1, K.psd_dmax16, in, projbuf);
    psd_launch_group<128>(h, K.d_psd.p + n1, n2 - n1, K.psd_dmax48, in, projbuf);
    psd_launch_group<512>(h, K.d_psd.p + n2, nc - n2, K.psd_max_d, in, projbuf);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(FOS_ERR_CUDA, std::string("PSD projection launch failed: ") + cudaGetErrorString(e));
}

}  // namespace fos
