// psd_large.cu -- K5 for large PSD cones (d > 112, up to 1024): projection onto the PSD cone,
// IndPSD(scaling=true) of ProximalOperators.jl as mapped at cones.jl:11 (dual through Moreau,
// cones.jl:80-85), for matrices that do not fit one SM's shared memory (config 4: d = 512).
//
// Algorithm (one cooperative group of CT CTAs per cone, all phases inside ONE kernel):
//   0. M = smat(x) with the diagonal multiplied by sqrt(2) (M = sqrt(2) X); sigma = (1+1/64)||M||_F.
//   1. G = M + sigma*I  (positive definite: its singular vectors ARE M's eigenvectors, also when M
//      has eigenvalues +lambda and -lambda, where a one-sided method on M itself is ill-posed).
//   2. One-sided (Hestenes) block Jacobi on the columns of G: G <- G J until the columns are mutually
//      orthogonal.  Columns are grouped in 2*CT blocks of bs; a round-robin tournament over blocks
//      gives each CTA one pair of blocks per step.  The CTA pulls its 2*bs columns into shared
//      memory with bulk copies (TMA engine, cp.async.bulk), one warp rotates one column pair per
//      round (bs rounds cover the bs^2 cross pairs), pushes the columns back, and the group meets at
//      a global-memory barrier.  Rotations only touch the two columns involved, so the bs pairs of a
//      round and the CT block pairs of a step are independent.
//   3. v_j = g_j/||g_j||, lambda_j = v_j' M v_j (Rayleigh quotient with the ORIGINAL matrix: the
//      shift costs ||M||_F/|lambda| in relative accuracy of ||g_j|| - sigma, the quotient does not),
//      w_j = sqrt(max(lambda_j,0)) v_j.
//   4. P = W W' (lower triangle, 64x64 register-tiled), repack, diag / sqrt(2), Moreau for the dual.
// Sweeps stop when the largest |cos| between two columns seen in a sweep is <= 1e-13, or <= 3e-8
// after falling by >= 100x in one sweep (quadratic phase: the next sweep would find ~1e-15).
// Deterministic: fixed pair order, fixed reduction trees, no floating-point atomics.
#include <algorithm>

#include "solver.cuh"

namespace fos {

constexpr int PL_MAX_SWEEPS = 48;
constexpr int PL_MAX_BLOCKS = 128;  // d_pad / BS for d <= 1024
constexpr int PL_TILE = 64;
constexpr int PL_JC = 32;  // columns of W per shared-memory chunk of the final product

struct PsdLargeCtl {  // one per cone, zeroed before every launch
    unsigned int bar;   // arrival counter of the group barrier
    int sweeps;
    unsigned int pad0_[30];
    unsigned int flag;  // last completed barrier target, on its own cache line (the waiters poll this one)
    unsigned int pad1_[31];
    unsigned long long maxcos[PL_MAX_SWEEPS];  // bit pattern of the largest |cos| seen in sweep k
    unsigned int ready[PL_MAX_BLOCKS];  // version of every column block in global memory: 1 after the initial
                                        // fill, +1 per Jacobi step (each block is rotated once per step)
};

struct PsdLargeArgs {
    const PsdCone *cones;
    const double *in;
    double *proj;
    double *work;
    int64_t work_stride;  // doubles per cone: G [d_pad][dS] | M [d_pad][dS] | lam [d_pad]
    PsdLargeCtl *ctl;
    double *vstore;       // per cone: the orthonormal eigenvector basis V of the previous projection [d_pad][dS]
    int64_t vstore_stride;
    int warm;             // 1: start from G = (M + sigma I) V_prev instead of G = M + sigma I
    int CT;     // CTAs per cone; the tournament has 2*CT blocks
    int d_pad;  // padded number of columns = 2*CT*bs
};

__device__ __forceinline__ void rr_pair_l(int s, int k, int D, int &a, int &b)
{
    const int M = D - 1;  // round-robin tournament: player D-1 is fixed, the others rotate
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

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// barrier among the CT CTAs of one cone: one atomic per CTA on the arrival counter; the last arriver
// publishes the target on a separate cache line, which the others poll with relaxed loads
__device__ __forceinline__ void group_barrier(PsdLargeCtl *ctl, unsigned int &epoch, unsigned int CT)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch++;
        const unsigned int target = epoch * CT;
        __threadfence();
        const unsigned int old = atomicAdd(&ctl->bar, 1u);
        if (old + 1u == target) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&ctl->flag), "r"(target) : "memory");
        } else {
            unsigned int v;
            do {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&ctl->flag) : "memory");
            } while (v < target);
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void bulk_store_1d(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
}

template <int DK>
struct PLCfg {
    static constexpr int BS = 8;                 // columns per block == warps per CTA
    static constexpr int THREADS = 32 * BS;
    static constexpr int DS = 64 * DK;           // padded column length
    static constexpr size_t COLS_BYTES = (size_t)2 * BS * DS * sizeof(double);
    static constexpr size_t PROD_BYTES = (size_t)2 * 32 * 64 * sizeof(double);  // Wi / Wk chunks of the final product
    static constexpr size_t SMEM = COLS_BYTES > PROD_BYTES ? COLS_BYTES : PROD_BYTES;
};

// one warp: orthogonalise columns P and Q (shared memory, DS doubles each).  n2p / n2q hold the squared
// column norms: computed when the block is loaded, updated here with the rotation identities
// alpha' = alpha - t*gamma, beta' = beta + t*gamma (at most 2*BS-1 updates before the next reload).
// lmax2 tracks the largest cos^2 seen.
template <int DK>
__device__ __forceinline__ void rotate_pair(double *__restrict__ cp, double *__restrict__ cq, double *n2p, double *n2q,
                                            int lane, double tiny2, double &lmax2)
{
    double2 P[DK], Q[DK];
    double g0 = 0.0, g1 = 0.0;
#pragma unroll
    for (int k = 0; k < DK; k++) {
        P[k] = *reinterpret_cast<const double2 *>(cp + 64 * k + 2 * lane);
        Q[k] = *reinterpret_cast<const double2 *>(cq + 64 * k + 2 * lane);
        g0 = fma(P[k].x, Q[k].x, g0);
        g1 = fma(P[k].y, Q[k].y, g1);
    }
    const double ga = warp_sum(g0 + g1);
    const double al = *n2p, be = *n2q;
    const double ab = al * be;
    if (!(ab > tiny2)) return;  // a zero (padding) column
    const double g2 = ga * ga;
    lmax2 = fmax(lmax2, g2 / ab);
    if (!(g2 > 1e-30 * ab)) return;  // |cos| <= 1e-15
    const double delta = be - al;
    const double hyp = sqrt(fma(delta, delta, 4.0 * g2));
    const double t = (delta >= 0.0 ? 2.0 : -2.0) * ga / (fabs(delta) + hyp);
    const double c = rsqrt(fma(t, t, 1.0));
    const double s = c * t;
#pragma unroll
    for (int k = 0; k < DK; k++) {
        double2 pn, qn;
        pn.x = fma(-s, Q[k].x, c * P[k].x);
        pn.y = fma(-s, Q[k].y, c * P[k].y);
        qn.x = fma(s, P[k].x, c * Q[k].x);
        qn.y = fma(s, P[k].y, c * Q[k].y);
        *reinterpret_cast<double2 *>(cp + 64 * k + 2 * lane) = pn;
        *reinterpret_cast<double2 *>(cq + 64 * k + 2 * lane) = qn;
    }
    __syncwarp();  // every lane has read *n2p / *n2q before lane 0 replaces them
    if (lane == 0) {
        *n2p = fma(-t, ga, al);
        *n2q = fma(t, ga, be);
    }
}

// Same rotation with column P held in REGISTERS by its warp for all BS rounds of a step (shared memory
// bandwidth, not FP64, bounds the sweep: this halves the LDS/STS traffic).  al = ||P||^2 travels with it.
template <int DK>
__device__ __forceinline__ void rotate_pair_reg(double2 (&P)[DK], double &al, double *__restrict__ cq, double *n2q,
                                                int lane, double tiny2, double &lmax2)
{
    double2 Q[DK];
    double g0 = 0.0, g1 = 0.0;
#pragma unroll
    for (int k = 0; k < DK; k++) {
        Q[k] = *reinterpret_cast<const double2 *>(cq + 64 * k + 2 * lane);
        g0 = fma(P[k].x, Q[k].x, g0);
        g1 = fma(P[k].y, Q[k].y, g1);
    }
    const double ga = warp_sum(g0 + g1);
    const double be = *n2q;
    const double ab = al * be;
    if (!(ab > tiny2)) return;
    const double g2 = ga * ga;
    lmax2 = fmax(lmax2, g2 / ab);
    if (!(g2 > 1e-30 * ab)) return;
    const double delta = be - al;
    const double hyp = sqrt(fma(delta, delta, 4.0 * g2));
    const double t = (delta >= 0.0 ? 2.0 : -2.0) * ga / (fabs(delta) + hyp);
    const double c = rsqrt(fma(t, t, 1.0));
    const double s = c * t;
#pragma unroll
    for (int k = 0; k < DK; k++) {
        double2 pn, qn;
        pn.x = fma(-s, Q[k].x, c * P[k].x);
        pn.y = fma(-s, Q[k].y, c * P[k].y);
        qn.x = fma(s, P[k].x, c * Q[k].x);
        qn.y = fma(s, P[k].y, c * Q[k].y);
        P[k] = pn;
        *reinterpret_cast<double2 *>(cq + 64 * k + 2 * lane) = qn;
    }
    __syncwarp();  // every lane has read *n2q before lane 0 replaces it
    if (lane == 0) *n2q = fma(t, ga, be);
    al = fma(-t, ga, al);
}

// acc[j] = sum_k M[i][k] * vc[j][k] for the BS columns vc (shared memory) and row i; M symmetric, column k
// contiguous in global memory (so the loads of a warp are coalesced over i)
template <int BS, int DS>
__device__ __forceinline__ void sym_times_cols(const double *__restrict__ M0, int d, const double *vc, int i,
                                               double (&acc)[BS])
{
#pragma unroll
    for (int j = 0; j < BS; j++) acc[j] = 0.0;
    for (int k = 0; k < d; k += 2) {
        const double a0 = M0[(size_t)k * DS + i];
        const double a1 = (k + 1 < d) ? M0[(size_t)(k + 1) * DS + i] : 0.0;
#pragma unroll
        for (int j = 0; j < BS; j++) {
            const double2 v = *reinterpret_cast<const double2 *>(vc + (size_t)j * DS + k);
            acc[j] = fma(a0, v.x, acc[j]);
            acc[j] = fma(a1, v.y, acc[j]);
        }
    }
}

template <int DK>
__global__ void __launch_bounds__(PLCfg<DK>::THREADS, 1) k5_psd_hestenes(const PsdLargeArgs a)
{
    using Cfg = PLCfg<DK>;
    constexpr int BS = Cfg::BS, T = Cfg::THREADS, DS = Cfg::DS;
    extern __shared__ __align__(128) unsigned char pl_smem[];
    double *cols = reinterpret_cast<double *>(pl_smem);  // [2*BS][DS]
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ double s_w[T / 32][2 * BS > 4 ? 2 * BS : 4];
    __shared__ double s_lam[2 * BS];
    __shared__ double s_n2[2 * BS];
    __shared__ int s_plist[1024 + 64];
    __shared__ int s_npos;

    const int cone = blockIdx.x / a.CT, cta = blockIdx.x - cone * a.CT;
    const PsdCone C = a.cones[cone];
    const int d = C.d;
    const int NBk = 2 * a.CT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *G = a.work + (size_t)cone * a.work_stride;
    double *M0 = G + (size_t)a.d_pad * DS;
    double *lam_g = M0 + (size_t)a.d_pad * DS;
    PsdLargeCtl *ctl = a.ctl + cone;
    unsigned int epoch = 0;
    const double sq2 = 1.4142135623730951;
    const double sgn = C.dual ? -1.0 : 1.0;  // dual: project -x, then add x (cones.jl:80-85)
    const int64_t plen = (int64_t)d * (d + 1) / 2;
    const double *xin = a.in + C.off;

    if (threadIdx.x == 0) {
        mbar_init(&s_mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t mphase = 0;

    // ---- phase 0: ||M||_F (every CTA of the group computes the identical value) ----
    double fro2;
    {
        double acc = 0.0;
        for (int64_t k = threadIdx.x; k < plen; k += T) acc = fma(xin[k], xin[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) s_w[warp][0] = acc;
        __syncthreads();
        double t = 0.0;
        for (int k = 0; k < T / 32; k++) t += s_w[k][0];
        fro2 = 2.0 * t;
        __syncthreads();
    }
    if (!(fro2 > 0.0) || !(fro2 < INFINITY)) {
        // zero matrix -> projection 0 (NaN/Inf input: propagate)
        const double fill = fro2 == 0.0 ? 0.0 : NAN;
        for (int64_t k = (int64_t)cta * T + threadIdx.x; k < plen; k += (int64_t)a.CT * T)
            a.proj[C.off + k] = C.dual ? __dadd_rn(xin[k], fill) : fill;
        return;
    }
    const double sigma = sqrt(fro2) * (1.0 + 1.0 / 64.0);
    const double tiny2 = (1e-32 * fro2) * (1e-32 * fro2);  // (1e-16 ||M||_F)^4 against alpha*beta

    // ---- phase 1: G = M + sigma I and a copy of M, columns of blocks 2*cta, 2*cta+1 ----
    {
        const int j0 = 2 * cta * BS;
        for (int idx = threadIdx.x; idx < 2 * BS * DS; idx += T) {
            const int j = j0 + idx / DS, i = idx % DS;
            double v = 0.0;
            if (i < d && j < d) {
                const int lo = i > j ? i : j, hi = i > j ? j : i;
                const int64_t k = (int64_t)hi * d - (int64_t)hi * (hi - 1) / 2 + (lo - hi);
                v = sgn * xin[k];
                if (i == j) v *= sq2;
            }
            M0[(size_t)j * DS + i] = v;
            G[(size_t)j * DS + i] = (i == j && i < d) ? v + sigma : v;
        }
    }
    group_barrier(ctl, epoch, a.CT);
    constexpr uint32_t BLK_BYTES0 = (uint32_t)(BS * DS * sizeof(double));
    double *Vst = a.vstore + (size_t)cone * a.vstore_stride;
    if (a.warm) {
        // Warm start: consecutive solver iterates have nearly the same eigenvectors, so start the sweeps from
        // G = (M + sigma I) V_prev (columns almost orthogonal already; V_prev is orthonormal, so the singular
        // vectors of G are still the eigenvectors of M).  Each CTA builds the columns of its two blocks.
        double *v0 = Vst + (size_t)(2 * cta) * BS * DS;
        double *g0w = G + (size_t)(2 * cta) * BS * DS;
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_expect_tx(&s_mbar, 2 * BLK_BYTES0);
            bulk_load_1d(cols, v0, 2 * BLK_BYTES0, &s_mbar);
        }
        mbar_wait(&s_mbar, mphase);
        mphase ^= 1;
        for (int half = 0; half < 2; half++) {
            const double *vc = cols + (size_t)half * BS * DS;
            for (int i = threadIdx.x; i < DS; i += T) {
                double acc[BS];
                if (i < d) sym_times_cols<BS, DS>(M0, d, vc, i, acc);
#pragma unroll
                for (int j = 0; j < BS; j++)
                    g0w[(size_t)(half * BS + j) * DS + i] = i < d ? fma(sigma, vc[(size_t)j * DS + i], acc[j]) : 0.0;
            }
        }
        group_barrier(ctl, epoch, a.CT);
    }

    // ---- phase 2: block Jacobi sweeps ----
    // Between steps a CTA only depends on the two CTAs that rotated its next blocks: every block carries a
    // version counter in global memory (ready[]), bumped by its owner after the store and awaited by the next
    // owner before the load.  Only the current owner reads or writes a block, so there is no WAR hazard and no
    // group-wide barrier inside a sweep (one per sweep remains, for the convergence test).
    constexpr uint32_t BLK_BYTES = (uint32_t)(BS * DS * sizeof(double));
    double prev_max = 1.0;
    int sweeps_done = 0;
    unsigned int gstep = 0;  // global step counter: blocks must have version gstep + 1
    for (int sweep = 0; sweep < PL_MAX_SWEEPS; sweep++) {
        double lmax = 0.0;
        for (int step = 0; step < NBk - 1; step++, gstep++) {
            int ba, bb;
            rr_pair_l(step, cta, NBk, ba, bb);
            double *ga_ = G + (size_t)ba * BS * DS, *gb_ = G + (size_t)bb * BS * DS;
            if (threadIdx.x == 0) {
                if (gstep > 0) {
                    const unsigned int want = gstep + 1u;
                    unsigned int va, vb;
                    do {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(va) : "l"(&ctl->ready[ba]) : "memory");
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(vb) : "l"(&ctl->ready[bb]) : "memory");
                    } while (va < want || vb < want);
                    __threadfence();
                }
                asm volatile("fence.proxy.async;" ::: "memory");
                mbar_expect_tx(&s_mbar, 2 * BLK_BYTES);
                bulk_load_1d(cols, ga_, BLK_BYTES, &s_mbar);
                bulk_load_1d(cols + BS * DS, gb_, BLK_BYTES, &s_mbar);
            }
            mbar_wait(&s_mbar, mphase);
            mphase ^= 1;
            for (int cidx = warp; cidx < 2 * BS; cidx += T / 32) {
                const double *cp = cols + (size_t)cidx * DS;
                double n0 = 0.0, n1 = 0.0;
#pragma unroll
                for (int k = 0; k < DK; k++) {
                    const double2 v = *reinterpret_cast<const double2 *>(cp + 64 * k + 2 * lane);
                    n0 = fma(v.x, v.x, n0);
                    n1 = fma(v.y, v.y, n1);
                }
                n0 = warp_sum(n0 + n1);
                if (lane == 0) s_n2[cidx] = n0;
            }
            __syncthreads();
            if (step == 0) {
                // pairs inside each of the two blocks, once per sweep
                for (int st = 0; st < BS - 1; st++) {
                    const int blk = warp / (BS / 2), k = warp - blk * (BS / 2);
                    int p, q;
                    rr_pair_l(st, k, BS, p, q);
                    rotate_pair<DK>(cols + (size_t)(blk * BS + p) * DS, cols + (size_t)(blk * BS + q) * DS,
                                    s_n2 + blk * BS + p, s_n2 + blk * BS + q, lane, tiny2, lmax);
                    __syncthreads();
                }
            }
            {
                // cross pairs: warp w keeps column w of the first block in registers through the BS rounds
                double2 P[DK];
                double *cp = cols + (size_t)warp * DS;
#pragma unroll
                for (int k = 0; k < DK; k++) P[k] = *reinterpret_cast<const double2 *>(cp + 64 * k + 2 * lane);
                double al = s_n2[warp];
                for (int r = 0; r < BS; r++) {
                    const int q = BS + ((warp + r) & (BS - 1));
                    rotate_pair_reg<DK>(P, al, cols + (size_t)q * DS, s_n2 + q, lane, tiny2, lmax);
                    __syncthreads();
                }
#pragma unroll
                for (int k = 0; k < DK; k++) *reinterpret_cast<double2 *>(cp + 64 * k + 2 * lane) = P[k];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_store_1d(ga_, cols, BLK_BYTES);
                bulk_store_1d(gb_, cols + BS * DS, BLK_BYTES);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&ctl->ready[ba]), "r"(gstep + 2u) : "memory");
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&ctl->ready[bb]), "r"(gstep + 2u) : "memory");
            }
            if (step == NBk - 2) {
                // publish this CTA's largest |cos| of the sweep (non-negative doubles order like integers)
                if (lane == 0) s_w[warp][0] = lmax;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double mx = 0.0;
                    for (int k = 0; k < T / 32; k++) mx = fmax(mx, s_w[k][0]);
                    mx = sqrt(mx);  // lmax tracks cos^2
                    atomicMax(&ctl->maxcos[sweep], (unsigned long long)__double_as_longlong(mx));
                }
                group_barrier(ctl, epoch, a.CT);  // once per sweep: everybody's maxcos is in
            } else {
                __syncthreads();  // thread 0 has finished the store before the buffer is reused
            }
        }
        sweeps_done = sweep + 1;
        const double mx = __longlong_as_double((long long)*((volatile unsigned long long *)&ctl->maxcos[sweep]));
        if (mx <= 1e-13 || (sweep > 0 && mx <= 3e-8 && mx <= 0.01 * prev_max)) break;
        prev_max = mx;
    }
    if (cta == 0 && threadIdx.x == 0) ctl->sweeps = sweeps_done;

    // ---- phase 3: own columns -> v_j, lambda_j = v_j' M v_j, w_j = sqrt(lambda_j+) v_j ----
    {
        double *g0 = G + (size_t)(2 * cta) * BS * DS;
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_expect_tx(&s_mbar, 2 * BLK_BYTES);
            bulk_load_1d(cols, g0, BLK_BYTES, &s_mbar);
            bulk_load_1d(cols + BS * DS, g0 + BS * DS, BLK_BYTES, &s_mbar);
        }
        mbar_wait(&s_mbar, mphase);
        mphase ^= 1;
        for (int cidx = warp; cidx < 2 * BS; cidx += T / 32) {
            double *cp = cols + (size_t)cidx * DS;
            double nn = 0.0;
            for (int i = lane; i < DS; i += 32) nn = fma(cp[i], cp[i], nn);
            nn = warp_sum(nn);
            const double inv = nn > 0.0 ? 1.0 / sqrt(nn) : 0.0;
            for (int i = lane; i < DS; i += 32) cp[i] *= inv;
        }
        // keep the orthonormal basis for the next projection of this cone (warm start)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store_1d(Vst + (size_t)(2 * cta) * BS * DS, cols, 2 * BLK_BYTES);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the source may be modified again
        }
        __syncthreads();
        // Y = M V_J by halves of BS columns (register budget), lambda_j = sum_i v_ij y_ij
        for (int half = 0; half < 2; half++) {
            const double *vc = cols + (size_t)half * BS * DS;
            double lp[BS];
#pragma unroll
            for (int j = 0; j < BS; j++) lp[j] = 0.0;
            for (int i = threadIdx.x; i < DS; i += T) {
                if (i >= d) continue;
                double acc[BS];
                sym_times_cols<BS, DS>(M0, d, vc, i, acc);
#pragma unroll
                for (int j = 0; j < BS; j++) lp[j] = fma(acc[j], vc[(size_t)j * DS + i], lp[j]);
            }
#pragma unroll
            for (int j = 0; j < BS; j++) {
                const double v = warp_sum(lp[j]);
                if (lane == 0) s_w[warp][half * BS + j] = v;
            }
        }
        __syncthreads();
        if (threadIdx.x < 2 * BS) {
            double t = 0.0;
            for (int k = 0; k < T / 32; k++) t += s_w[k][threadIdx.x];
            s_lam[threadIdx.x] = t;
            lam_g[2 * cta * BS + threadIdx.x] = t;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 2 * BS * DS; idx += T) {
            const double l = s_lam[idx / DS];
            cols[idx] = l > 0.0 ? cols[idx] * sqrt(l) : 0.0;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store_1d(g0, cols, 2 * BLK_BYTES);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        group_barrier(ctl, epoch, a.CT);
    }

    // ---- phase 4: P = W W' over the columns with lambda > 0 ----
    if (warp == 0) {
        int cnt = 0;
        for (int base = 0; base < a.d_pad; base += 32) {
            const bool pos = base + lane < a.d_pad && lam_g[base + lane] > 0.0;
            const unsigned int m = __ballot_sync(0xffffffffu, pos);
            if (pos) s_plist[cnt + __popc(m & ((1u << lane) - 1u))] = base + lane;
            cnt += __popc(m);
        }
        if (lane == 0) s_npos = cnt;
    }
    __syncthreads();
    const int npos = s_npos;
    const int nT = (d + PL_TILE - 1) / PL_TILE;
    const int ntiles = nT * (nT + 1) / 2;
    double *Wi = cols, *Wk = cols + PL_JC * PL_TILE;
    const int ti = threadIdx.x & 15, tk = (threadIdx.x >> 4) & 15;
    const bool computes = threadIdx.x < 256;
    for (int t = cta; t < ntiles; t += a.CT) {
        // tile index -> (I, K), I >= K
        int I = (int)floor((sqrt(8.0 * t + 1.0) - 1.0) / 2.0);
        while (I * (I + 1) / 2 > t) I--;
        while ((I + 1) * (I + 2) / 2 <= t) I++;
        const int K = t - I * (I + 1) / 2;
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) acc[x][y] = 0.0;
        for (int c0 = 0; c0 < npos; c0 += PL_JC) {
            const int nc = npos - c0 < PL_JC ? npos - c0 : PL_JC;
            __syncthreads();
            for (int idx = threadIdx.x; idx < PL_JC * PL_TILE; idx += T) {
                const int jj = idx / PL_TILE, r = idx - jj * PL_TILE;
                double vi = 0.0, vk = 0.0;
                if (jj < nc) {
                    const double *col = G + (size_t)s_plist[c0 + jj] * DS;
                    vi = col[I * PL_TILE + r];
                    vk = col[K * PL_TILE + r];
                }
                Wi[idx] = vi;
                Wk[idx] = vk;
            }
            __syncthreads();
            if (computes) {
#pragma unroll 4
                for (int jj = 0; jj < PL_JC; jj++) {
                    const double2 i01 = *reinterpret_cast<const double2 *>(Wi + jj * PL_TILE + 4 * ti);
                    const double2 i23 = *reinterpret_cast<const double2 *>(Wi + jj * PL_TILE + 4 * ti + 2);
                    const double2 k01 = *reinterpret_cast<const double2 *>(Wk + jj * PL_TILE + 4 * tk);
                    const double2 k23 = *reinterpret_cast<const double2 *>(Wk + jj * PL_TILE + 4 * tk + 2);
                    const double wi[4] = {i01.x, i01.y, i23.x, i23.y};
                    const double wk[4] = {k01.x, k01.y, k23.x, k23.y};
#pragma unroll
                    for (int x = 0; x < 4; x++)
#pragma unroll
                        for (int y = 0; y < 4; y++) acc[x][y] = fma(wi[x], wk[y], acc[x][y]);
                }
            }
        }
        if (computes) {
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const int i = I * PL_TILE + 4 * ti + x, k = K * PL_TILE + 4 * tk + y;
                    if (i < d && k <= i) {
                        const int64_t idx = (int64_t)k * d - (int64_t)k * (k - 1) / 2 + (i - k);
                        double v = acc[x][y];
                        if (i == k) v /= sq2;
                        a.proj[C.off + idx] = C.dual ? __dadd_rn(xin[idx], v) : v;
                    }
                }
        }
    }
}

// V = I for every cone: an orthonormal basis to start from
__global__ void k_psd_identity(double *V, int nc, int d_pad, int dS)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nc * d_pad) return;
    const int cone = idx / d_pad, j = idx - cone * d_pad;
    if (j < dS) V[((size_t)cone * d_pad + j) * dS + j] = 1.0;
}

template <int DK>
static void launch_large(Handle *h, const PsdLargeArgs &a, int ncones)
{
    using Cfg = PLCfg<DK>;
    FOS_CUDA(cudaFuncSetAttribute(k5_psd_hestenes<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    void *params[] = {(void *)&a};
    FOS_CUDA(cudaLaunchCooperativeKernel((void *)k5_psd_hestenes<DK>, dim3((unsigned)(ncones * a.CT)), dim3(Cfg::THREADS),
                                         params, Cfg::SMEM, h->stream));
    h->stats.launches++;
}

// Projects the cones K.psd_large[...] (d > PSD_SMEM_MAX_D).  Cones are processed in chunks that fit the
// GPU as one co-resident cooperative grid (CT CTAs per cone).
void psd_project_large(Handle *h, ConeSet &K, const double *in, double *projbuf)
{
    const int nc = (int)K.psd_large.size();
    if (nc == 0) return;
    int dmax = 0;
    for (const PsdCone &c : K.psd_large) dmax = std::max(dmax, (int)c.d);
    static const int dks[] = {2, 4, 6, 8, 12, 16};
    int DK = 0;
    for (int k : dks)
        if (64 * k >= dmax) {
            DK = k;
            break;
        }
    if (DK == 0)
        throw Error(FOS_ERR_UNSUPPORTED, "SDP cone of order " + std::to_string(dmax) + " exceeds the supported 1024");
    const int bs = 8;
    const int dS = 64 * DK;
    PsdLargeArgs a;
    a.d_pad = (int)ru(dmax, 2 * bs);
    a.CT = a.d_pad / (2 * bs);
    a.work_stride = ru((int64_t)2 * a.d_pad * dS + a.d_pad, 16);
    const int per_launch = std::max(1, h->num_sms / a.CT);
    const int chunk = std::min(per_launch, nc);
    if (K.psd_work.n < (size_t)a.work_stride * chunk) K.psd_work.alloc((size_t)a.work_stride * chunk);
    if (K.psd_ctl.n < (size_t)chunk * sizeof(PsdLargeCtl)) K.psd_ctl.alloc((size_t)chunk * sizeof(PsdLargeCtl));
    // eigenvector bases of the previous call, one per cone (warm start); cold when the geometry changed
    const size_t vneed = (size_t)a.d_pad * dS * nc;
    if (K.psd_vstore.n != vneed) {
        K.psd_vstore.alloc(vneed);  // zeroed
        k_psd_identity<<<(unsigned)(nc * a.d_pad + 255) / 256, 256, 0, h->stream>>>(K.psd_vstore.p, nc, a.d_pad, dS);
        K.psd_warm = true;  // V = I is a valid basis: the first projection is the cold start
    }
    a.vstore_stride = (int64_t)a.d_pad * dS;
    a.in = in;
    a.proj = projbuf;
    a.work = K.psd_work.p;
    a.ctl = reinterpret_cast<PsdLargeCtl *>(K.psd_ctl.p);
    for (int c0 = 0; c0 < nc; c0 += chunk) {
        const int n = std::min(chunk, nc - c0);
        a.cones = K.d_psd_large.p + c0;
        a.vstore = K.psd_vstore.p + (size_t)c0 * a.vstore_stride;
        a.warm = (K.psd_warm && K.psd_warm_enabled) ? 1 : 0;
        FOS_CUDA(cudaMemsetAsync(K.psd_ctl.p, 0, (size_t)n * sizeof(PsdLargeCtl), h->stream));
        switch (DK) {
        case 2: launch_large<2>(h, a, n); break;
        case 4: launch_large<4>(h, a, n); break;
        case 6: launch_large<6>(h, a, n); break;
        case 8: launch_large<8>(h, a, n); break;
        case 12: launch_large<12>(h, a, n); break;
        default: launch_large<16>(h, a, n); break;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw Error(FOS_ERR_CUDA, std::string("large PSD projection launch failed: ") + cudaGetErrorString(e));
}

// number of Jacobi sweeps the last large projection needed for cone 0 of its last chunk (diagnostics)
int psd_large_last_sweeps(Handle *h, ConeSet &K)
{
    if (K.psd_ctl.n == 0) return 0;
    PsdLargeCtl c;
    FOS_CUDA(cudaStreamSynchronize(h->stream));
    FOS_CUDA(cudaMemcpy(&c, K.psd_ctl.p, sizeof(c), cudaMemcpyDeviceToHost));
    return c.sweeps;
}

}  // namespace fos
