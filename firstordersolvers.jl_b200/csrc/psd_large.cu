// psd_large.cu -- K5 for large PSD cones (d > 112, up to 1024): projection onto the PSD cone,
// IndPSD(scaling=true) of ProximalOperators.jl as mapped at cones.jl:11 (dual through Moreau,
// cones.jl:80-85), for matrices that do not fit one SM's shared memory (config 4: d = 512).
//
// Algorithm (one cooperative group of CT CTAs per cone, all phases inside ONE kernel):
//   0. M = smat(x) with the diagonal multiplied by sqrt(2) (M = sqrt(2) X); sigma = (1+1/64)||M||_F.
//   1. G = M + sigma*I  (positive definite: its singular vectors ARE M's eigenvectors, also when M
//      has eigenvalues +lambda and -lambda, where a one-sided method on M itself is ill-posed).
//   2. One-sided (Hestenes) block Jacobi on the columns of G: G <- G J until the columns are mutually
//      orthogonal.  Columns are grouped in 2*CT blocks of 8; a round-robin tournament over blocks gives each CTA
//      one pair of blocks per step.  Per step the CTA
//        a. reads its 16 columns from global memory, recognising the other CTAs' output by a version bit in the
//           data itself (no flags, no fences: see "block exchange" below),
//        b. forms their 16 x 16 Gram block S = X'X on the FP64 tensor cores (DMMA m8n8k4),
//        c. runs the step's 8 rounds of 8 disjoint rotations ON S (S <- R'SR, J <- JR: the same rotations, in
//           exact arithmetic, as rotating the length-d columns pair by pair, without the length-d dot product
//           and warp reduction in every round's dependency chain; G is well conditioned by construction, so
//           working from the Gram block costs no accuracy),
//        d. applies X <- X J on the tensor cores and stores the result straight to the next version's location.
//      Rotations only touch the two columns involved, so the 8 pairs of a round and the CT block pairs of a
//      step are independent.  Measured per step at d = 512 (SM cycles): exchange 2950, Gram 2500, rotations
//      4000, X J 3650 (profiles/r2_psd_large.md); the kernel-per-pair version this replaces spent 4.8 us of an
//      11.3 us step on the flag-based exchange alone.
//   3. v_j = g_j/||g_j||, lambda_j = v_j' M v_j (Rayleigh quotient with the ORIGINAL matrix: the
//      shift costs ||M||_F/|lambda| in relative accuracy of ||g_j|| - sigma, the quotient does not),
//      w_j = sqrt(max(lambda_j,0)) v_j.
//   4. P = W W' (lower triangle, 64x64 register-tiled), repack, diag / sqrt(2), Moreau for the dual.
// Sweeps stop when the largest |cos| between two columns seen in a sweep is <= 1e-13, or <= 3e-8
// after falling by >= 100x in one sweep (quadratic phase: the next sweep would find ~1e-15).
// Deterministic: fixed pair order, fixed reduction trees, no floating-point atomics.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "solver.cuh"

namespace fos {

constexpr int PL_MAX_SWEEPS = 48;
constexpr unsigned int PL_SPIN_LIMIT = 1u << 22;  // polls of one thread for one block (seconds) before the kernel gives up
constexpr int PL_NBUF = 8;         // G buffers: version v of a block lives in buffer v % PL_NBUF ...
constexpr int PL_SYNC_STEPS = 8;   // ... and the CTAs of a cone meet every PL_SYNC_STEPS steps (see "block exchange" below)
constexpr int PL_TILE = 64;
constexpr int PL_JC = 32;  // columns of W per shared-memory chunk of the final product

struct PsdLargeCtl {  // one per cone, zeroed before every launch
    unsigned int bar;   // arrival counter of the group barrier
    int sweeps;
    unsigned int pad0_[30];
    unsigned int flag;  // last completed barrier target, on its own cache line (the waiters poll this one)
    unsigned int pad1_[31];
    unsigned long long maxcos[PL_MAX_SWEEPS];  // bit pattern of the largest |cos| seen in sweep k
};

struct PsdLargeArgs {
    const PsdCone *cones;
    const double *in;
    double *proj;
    double *work;
    int64_t work_stride;  // doubles per cone: PL_NBUF x G [d_pad][dS] | M [d_pad][dS] | lam [d_pad]
    PsdLargeCtl *ctl;
    double *vstore;       // per cone: the orthonormal eigenvector basis V of the previous projection [d_pad][dS]
    int64_t vstore_stride;
    int warm;             // 1: start from G = (M + sigma I) V_prev instead of G = M + sigma I
    unsigned long long *prof;  // diagnostics (FOS_PSD_PROF=1): SM clock cycles per part of the step, summed over CTAs
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
    static constexpr int ROWS = 64 * DK;         // rows that take part in the arithmetic (d <= ROWS, the rest zero)
    static constexpr int DS = ROWS + 4;          // column stride: +4 doubles keeps the MMA fragment loads of the Gram
                                                 // step (8 columns x 4 rows per warp) free of bank conflicts
    static constexpr size_t COLS_BYTES = (size_t)2 * BS * DS * sizeof(double);
    static constexpr size_t PROD_BYTES = (size_t)2 * 32 * 64 * sizeof(double);  // Wi / Wk chunks of the final product
    static constexpr size_t SMEM = COLS_BYTES > PROD_BYTES ? COLS_BYTES : PROD_BYTES;
};

// D(8x8) += A(8x4) B(4x8) on the FP64 tensor cores.  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c = {D[lane/4][2*(lane%4)], D[lane/4][2*(lane%4)+1]}.
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// 1/sqrt(w) to full precision for a normal positive w, without the branches of rsqrt(): hardware seed (about 20
// bits) and two Newton steps.  w = 0 gives NaN (the callers select it away).
__device__ __forceinline__ double rsqrt_nr(double w)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
    double e = fma(-(w * y), y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-(w * y), y, 1.0);
    return fma(0.5 * y, e, y);
}

// The Jacobi rotation that orthogonalises columns p < q, from their Gram entries al = S[p][p], be = S[q][q],
// g = S[p][q]:  p' = c p - s q,  q' = s p + c q  with tan(2 theta) = 2 g / (be - al), |theta| <= pi/4.
// Returns, for index i of the pair (partner ip), the coefficients of  x_i' = c x_i + e x_ip  and cos^2 between the
// two columns (20 bits are plenty for the convergence test).  Two rsqrt, no division, no branch:
// r1 = 1/sqrt(delta^2 + 4 g^2),  c^2 = (1 + |delta| r1)/2,  s = sign(delta) g r1 / c.
template <int LD>
__device__ __forceinline__ void jacobi_cs(const double (*S)[LD], int i, int ip, double tiny2, double &c, double &e,
                                          double &cos2)
{
    const int p = i < ip ? i : ip, q = i < ip ? ip : i;
    const double al = S[p][p], be = S[q][q], g = S[p][q];
    const double ab = al * be, g2 = g * g;
    double rab;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rab) : "d"(ab));
    const bool live = ab > tiny2;                // not a zero (padding) column
    const bool rot = live && g2 > 1e-30 * ab;    // |cos| > 1e-15
    const double delta = be - al;
    const double r1 = rsqrt_nr(fma(delta, delta, 4.0 * g2));
    const double cc = fma(0.5 * fabs(delta), r1, 0.5);
    const double r2 = rsqrt_nr(cc);
    const double sn = (delta >= 0.0 ? g : -g) * (r1 * r2);
    c = rot ? cc * r2 : 1.0;
    e = rot ? (i < ip ? -sn : sn) : 0.0;
    cos2 = live ? g2 * rab : 0.0;
}

// ---- block exchange between the CTAs of a cone: self-validating words, no flags and no fences -------------------
// Version v of a column block (v = number of steps it has been through) lives in G buffer v % PL_NBUF, and every
// double of it carries bit (v / PL_NBUF) & 1 in its last mantissa bit (a perturbation of one ulp, the size of the
// rounding of the rotation itself).  A reader polls the data, not a flag, and the writer's stores need no fence.
// What can a location hold when its reader polls for version v?  Older versions of the same block, v - PL_NBUF,
// v - 2 PL_NBUF, ...  The first carries the other bit; the second carries the SAME bit, so it must be impossible:
// the CTAs of a cone meet at a barrier every PL_SYNC_STEPS steps, every version up to the last barrier's step is
// written, and the reader is at most PL_SYNC_STEPS - 1 steps past it: the stalest content is younger than
// v - PL_SYNC_STEPS - PL_NBUF + 1 > v - 2 PL_NBUF.  (Without the barrier the lag is only bounded by the number of
// CTAs: a CTA may run one step ahead of its neighbour, two ahead of the next, ...; measured: with 64 CTAs and two
// buffers, stale blocks WERE accepted.)  Overwriting is safe for any PL_NBUF >= 2: version v + PL_NBUF is written by
// the CTA that read v + PL_NBUF - 1, which exists only after v was read.
static_assert(PL_SYNC_STEPS <= PL_NBUF, "the stalest possible content must carry the other version bit");
__device__ __forceinline__ unsigned long long version_bit(unsigned int ver) { return (ver / PL_NBUF) & 1u; }
__device__ __forceinline__ double tag_word(double v, unsigned long long bit)
{
    return __longlong_as_double((long long)(((unsigned long long)__double_as_longlong(v) & ~1ull) | bit));
}
__device__ __forceinline__ double2 ld_relaxed_v2(const double *p)
{
    double2 v;
    asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v2(double *p, double x, double y)
{
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ bool tag_ok(double2 v, unsigned long long bit)
{
    return (((unsigned long long)__double_as_longlong(v.x) & 1ull) == bit) &&
           (((unsigned long long)__double_as_longlong(v.y) & 1ull) == bit);
}
// rows 0..ROWS-1 of the BS columns of NB blocks (1 or 2), version `ver`, into shared memory (the 4 padding rows of a
// column stay zero).  Units that still carry the old bit are read again until they do not.
template <int DK, int BS, int T, int NB>
__device__ __forceinline__ void load_blocks_tagged(double *sdst, const double *gsrc0, const double *gsrc1, unsigned int ver)
{
    constexpr int ROWS = 64 * DK, DS = ROWS + 4, H = ROWS / 2, NU = BS * H / T;
    static_assert(BS * H % T == 0, "whole number of 16-byte units per thread");
    const unsigned long long bit = version_bit(ver);
    int off[NU];
#pragma unroll
    for (int k = 0; k < NU; k++) {
        const int u = threadIdx.x + T * k, col = u / H;
        off[k] = col * DS + 2 * (u - col * H);
    }
    double2 v[NB][NU];
    unsigned int spins = 0;  // a block that never arrives is a bug: stop the kernel instead of hanging the GPU
    // The writers store their rows in ascending order as well: poll the LAST unit of every block (one round trip per
    // poll for all blocks) and fetch the rest when it is there -- fetching everything up front was measured slower
    // (the blocks are normally still on their way when the reader starts, and the stale reads load the L2).
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int b = 0; b < NB; b++) v[b][NU - 1] = ld_relaxed_v2((b ? gsrc1 : gsrc0) + off[NU - 1]);
#pragma unroll
        for (int b = 0; b < NB; b++) ok = ok && tag_ok(v[b][NU - 1], bit);
        if (ok) break;
        if (++spins > PL_SPIN_LIMIT) {
            printf("k5_psd_hestenes: CTA %d thread %d waited for version %u: %016llx\n", (int)blockIdx.x, (int)threadIdx.x,
                   ver, (unsigned long long)__double_as_longlong(v[0][NU - 1].x));
            __trap();
        }
    }
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
        for (int k = 0; k < NU - 1; k++) v[b][k] = ld_relaxed_v2((b ? gsrc1 : gsrc0) + off[k]);
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
        for (int k = 0; k < NU; k++) {
            while (!tag_ok(v[b][k], bit)) {
                v[b][k] = ld_relaxed_v2((b ? gsrc1 : gsrc0) + off[k]);
                if (++spins > PL_SPIN_LIMIT) {
                    printf("k5_psd_hestenes: CTA %d thread %d waited for version %u of block %d unit %d: %016llx %016llx\n",
                           (int)blockIdx.x, (int)threadIdx.x, ver, b, k, (unsigned long long)__double_as_longlong(v[b][k].x),
                           (unsigned long long)__double_as_longlong(v[b][k].y));
                    __trap();
                }
            }
            *reinterpret_cast<double2 *>(sdst + (size_t)b * BS * DS + off[k]) = v[b][k];
        }
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
    constexpr int BS = Cfg::BS, T = Cfg::THREADS, DS = Cfg::DS, ROWS = Cfg::ROWS, NC = 2 * BS, NW = T / 32;
    static_assert(T == NC * NC && BS == 8, "one thread per entry of the NC x NC Gram block; 8x8 MMA tiles");
    extern __shared__ __align__(128) unsigned char pl_smem[];
    double *cols = reinterpret_cast<double *>(pl_smem);  // [2*BS][DS]
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ double s_w[T / 32][2 * BS > 4 ? 2 * BS : 4];
    __shared__ double s_lam[2 * BS];
    __shared__ __align__(16) double s_part[NW][3][64];  // per-warp partial Gram tiles: (a,a), (b,a), (b,b)
    __shared__ double s_S[2][NC][NC + 1];               // Gram block of the step's 2*BS columns, double-buffered
    __shared__ __align__(16) double s_J[2][NC][NC];     // accumulated rotations
    __shared__ signed char s_partner[2 * BS - 1][NC];   // rounds 0..BS-2: pairs inside each block; BS-1..2BS-2: cross pairs
    __shared__ int s_plist[1024 + 64];
    __shared__ int s_npos;

    const int cone = blockIdx.x / a.CT, cta = blockIdx.x - cone * a.CT;
    const PsdCone C = a.cones[cone];
    const int d = C.d;
    const int NBk = 2 * a.CT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *G = a.work + (size_t)cone * a.work_stride;  // PL_NBUF G buffers, M, lambda
    const size_t gbuf = (size_t)a.d_pad * DS;
    double *M0 = G + PL_NBUF * gbuf;
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
    for (int idx = threadIdx.x; idx < (2 * BS - 1) * NC; idx += T) {
        const int r = idx / NC, i = idx - r * NC;
        int partner = i;
        if (r < BS - 1) {
            const int blk = i / BS, li = i - blk * BS;
            for (int k = 0; k < BS / 2; k++) {
                int p, q;
                rr_pair_l(r, k, BS, p, q);
                if (p == li) partner = blk * BS + q;
                if (q == li) partner = blk * BS + p;
            }
        } else {
            const int rc = r - (BS - 1);
            partner = i < BS ? BS + ((i + rc) & (BS - 1)) : ((i - BS - rc) & (BS - 1));
        }
        s_partner[r][i] = (signed char)partner;
    }
    for (int idx = threadIdx.x; idx < 2 * BS * DS; idx += T) cols[idx] = 0.0;  // the block loads skip the padding rows

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
            G[(size_t)j * DS + i] = tag_word((i == j && i < d) ? v + sigma : v, 0ull);  // version 0
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
            bulk_load_1d(cols, v0, BLK_BYTES0, &s_mbar);  // one copy per block: a block is < 128 KB for every d <= 1024
            bulk_load_1d(cols + BS * DS, v0 + (size_t)BS * DS, BLK_BYTES0, &s_mbar);
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
                    g0w[(size_t)(half * BS + j) * DS + i] =
                        tag_word(i < d ? fma(sigma, vc[(size_t)j * DS + i], acc[j]) : 0.0, 0ull);
            }
        }
        group_barrier(ctl, epoch, a.CT);
    }

    // ---- phase 2: block Jacobi sweeps ----
    // Between steps a CTA only depends on the two CTAs that rotated its next blocks, and it recognises their
    // output by the version bit of the data itself (load_block_tagged): no flags and no fences; a group-wide barrier
    // every PL_SYNC_STEPS steps bounds the lag between CTAs, and one per sweep serves the convergence test.
    constexpr uint32_t BLK_BYTES = (uint32_t)(BS * DS * sizeof(double));
    double prev_max = 1.0;
    int sweeps_done = 0;
    unsigned int gstep = 0;  // steps done so far = version of every block
    for (int sweep = 0; sweep < PL_MAX_SWEEPS; sweep++) {
        double lmax = 0.0;
        for (int step = 0; step < NBk - 1; step++, gstep++) {
            int ba, bb;
            rr_pair_l(step, cta, NBk, ba, bb);
            long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0;
            if (a.prof) tp0 = clock64();
            {
                const double *src = G + (gstep % PL_NBUF) * gbuf;
                if constexpr (DK <= 8) {  // both blocks in flight at once while the registers allow it
                    load_blocks_tagged<DK, BS, T, 2>(cols, src + (size_t)ba * BS * DS, src + (size_t)bb * BS * DS, gstep);
                } else {
                    load_blocks_tagged<DK, BS, T, 1>(cols, src + (size_t)ba * BS * DS, nullptr, gstep);
                    load_blocks_tagged<DK, BS, T, 1>(cols + BS * DS, src + (size_t)bb * BS * DS, nullptr, gstep);
                }
            }
            __syncthreads();
            if (a.prof) tp1 = clock64();
            // S = X'X for the step's 2*BS columns on the FP64 tensor cores: warp w takes rows 8w..8w+7 of every
            // 64 (two k-steps, two accumulator sets); the partial tiles are summed across warps in a fixed order.
            {
                const int g = lane >> 2, k4 = lane & 3;
                const double *x0 = cols + (size_t)g * DS + k4;
                const double *x1 = cols + (size_t)(BS + g) * DS + k4;
                double c00[2] = {0.0, 0.0}, c10[2] = {0.0, 0.0}, c11[2] = {0.0, 0.0};
                double e00[2] = {0.0, 0.0}, e10[2] = {0.0, 0.0}, e11[2] = {0.0, 0.0};
#pragma unroll 2
                for (int i0 = 8 * warp; i0 < ROWS; i0 += 8 * NW) {
                    const double a0 = x0[i0], a1 = x1[i0], b0 = x0[i0 + 4], b1 = x1[i0 + 4];
                    dmma884(c00, a0, a0);
                    dmma884(c10, a1, a0);
                    dmma884(c11, a1, a1);
                    dmma884(e00, b0, b0);
                    dmma884(e10, b1, b0);
                    dmma884(e11, b1, b1);
                }
                double2 *pp = reinterpret_cast<double2 *>(&s_part[warp][0][g * 8 + 2 * k4]);
                pp[0] = make_double2(c00[0] + e00[0], c00[1] + e00[1]);
                pp[32] = make_double2(c10[0] + e10[0], c10[1] + e10[1]);
                pp[64] = make_double2(c11[0] + e11[0], c11[1] + e11[1]);
            }
            __syncthreads();
            const int ei_ = threadIdx.x >> 4, ej_ = threadIdx.x & (NC - 1);  // this thread's entry of S and J
            {
                int tile, idx;
                if (ei_ < BS) {
                    tile = ej_ < BS ? 0 : 1;
                    idx = ej_ < BS ? ei_ * 8 + ej_ : (ej_ - BS) * 8 + ei_;
                } else {
                    tile = ej_ < BS ? 1 : 2;
                    idx = ej_ < BS ? (ei_ - BS) * 8 + ej_ : (ei_ - BS) * 8 + (ej_ - BS);
                }
                double sv = 0.0;
#pragma unroll
                for (int w = 0; w < NW; w++) sv += s_part[w][tile][idx];
                s_S[0][ei_][ej_] = sv;
                s_J[0][ei_][ej_] = ei_ == ej_ ? 1.0 : 0.0;
            }
            __syncthreads();
            if (a.prof) tp2 = clock64();
            // The rotations of the step on the Gram block: S <- R'SR, J <- JR per round of NC/2 disjoint pairs (in exact
            // arithmetic what rotating the columns pair by pair does).  Every thread derives the rotation of its COLUMN
            // pair itself; the rotation of its row pair is the one lane ei_ of the same warp derived (lanes 0..15 hold
            // the column indices 0..15), so a round costs one barrier and one rotation per thread.
            int cur = 0;
            auto jacobi_round = [&](int ip, int jp) {
                double cj, ej, cos2j;
                jacobi_cs<NC + 1>(s_S[cur], ej_, jp, tiny2, cj, ej, cos2j);
                const double ci = __shfl_sync(0xffffffffu, cj, ei_);
                const double ei = __shfl_sync(0xffffffffu, ej, ei_);
                if (ei_ == jp && ej_ < jp) lmax = fmax(lmax, cos2j);  // one thread per pair
                const double sij = s_S[cur][ei_][ej_], sijp = s_S[cur][ei_][jp];
                const double sipj = s_S[cur][ip][ej_], sipjp = s_S[cur][ip][jp];
                s_S[cur ^ 1][ei_][ej_] = ci * fma(ej, sijp, cj * sij) + ei * fma(ej, sipjp, cj * sipj);
                s_J[cur ^ 1][ei_][ej_] = fma(ej, s_J[cur][ei_][jp], cj * s_J[cur][ei_][ej_]);
                cur ^= 1;
                __syncthreads();
            };
            if (step == 0)  // pairs inside each of the two blocks, once per sweep
                for (int r = 0; r < BS - 1; r++) jacobi_round(s_partner[r][ei_], s_partner[r][ej_]);
#pragma unroll
            for (int rc = 0; rc < BS; rc++) {  // cross pairs: column w of the first block meets column (w + rc) % BS of the second
                const int ip = ei_ < BS ? BS + ((ei_ + rc) & (BS - 1)) : ((ei_ - BS - rc) & (BS - 1));
                const int jp = ej_ < BS ? BS + ((ej_ + rc) & (BS - 1)) : ((ej_ - BS - rc) & (BS - 1));
                jacobi_round(ip, jp);
            }
            if (a.prof) tp3 = clock64();
            // X <- X J as Y' = J'X' on the FP64 tensor cores (the J fragments stay in registers, every element of X is
            // read once), straight into the next version's location with the version bit set: fire and forget.
            // Warp w takes rows 8w..8w+7 of every 64.
            {
                const int g = lane >> 2, k4 = lane & 3;
                double ja[2][4];  // A fragments: J'[8 mt + g][4 ks + k4]
#pragma unroll
                for (int mt = 0; mt < 2; mt++)
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) ja[mt][ks] = s_J[cur][4 * ks + k4][8 * mt + g];
                double *dst = G + ((gstep + 1u) % PL_NBUF) * gbuf;
                double *da = dst + (size_t)ba * BS * DS + (size_t)g * DS + 2 * k4;
                double *db = dst + (size_t)bb * BS * DS + (size_t)g * DS + 2 * k4;
                const unsigned long long bit = version_bit(gstep + 1u);
                const double *xb = cols + (size_t)k4 * DS + g;  // B fragments: X[n0 + g][4 ks + k4]
#pragma unroll 2
                for (int n0 = 8 * warp; n0 < ROWS; n0 += 8 * NW) {
                    double ya[2] = {0.0, 0.0}, yb[2] = {0.0, 0.0};
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const double xv = xb[(size_t)(4 * ks) * DS + n0];
                        dmma884(ya, ja[0][ks], xv);
                        dmma884(yb, ja[1][ks], xv);
                    }
                    st_relaxed_v2(da + n0, tag_word(ya[0], bit), tag_word(ya[1], bit));  // rows n0 + 2 k4, + 1 of column g
                    st_relaxed_v2(db + n0, tag_word(yb[0], bit), tag_word(yb[1], bit));
                }
            }
            if (a.prof) tp4 = clock64();
            if (step == NBk - 2) {
                // publish this CTA's largest |cos| of the sweep (non-negative doubles order like integers)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
                if (lane == 0) s_w[warp][0] = lmax;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double mx = 0.0;
                    for (int k = 0; k < T / 32; k++) mx = fmax(mx, s_w[k][0]);
                    mx = sqrt(mx);  // lmax tracks cos^2
                    atomicMax(&ctl->maxcos[sweep], (unsigned long long)__double_as_longlong(mx));
                }
                group_barrier(ctl, epoch, a.CT);  // once per sweep: everybody's maxcos is in
            } else if ((step % PL_SYNC_STEPS) == PL_SYNC_STEPS - 1) {
                group_barrier(ctl, epoch, a.CT);  // bounds how far one CTA can run ahead of another (block exchange)
            } else {
                __syncthreads();  // everybody has read the columns before the next step's loads replace them
            }
            if (a.prof && threadIdx.x == 0) {
                atomicAdd(a.prof + 0, (unsigned long long)(tp1 - tp0));  // wait for + load the two blocks
                atomicAdd(a.prof + 1, (unsigned long long)(tp2 - tp1));  // Gram block
                atomicAdd(a.prof + 2, (unsigned long long)(tp3 - tp2));  // rotations
                atomicAdd(a.prof + 3, (unsigned long long)(tp4 - tp3));  // X J + stores
                atomicAdd(a.prof + 4, (unsigned long long)(clock64() - tp4));  // end-of-step barrier
                atomicAdd(a.prof + 5, 1ull);
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
        {   // the final version of this CTA's own two blocks (validated like every other block load)
            const double *src = G + (gstep % PL_NBUF) * gbuf + (size_t)(2 * cta) * BS * DS;
            load_blocks_tagged<DK, BS, T, 1>(cols, src, nullptr, gstep);
            load_blocks_tagged<DK, BS, T, 1>(cols + BS * DS, src + (size_t)BS * DS, nullptr, gstep);
        }
        __syncthreads();
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
            bulk_store_1d(Vst + (size_t)(2 * cta) * BS * DS, cols, BLK_BYTES);
            bulk_store_1d(Vst + (size_t)(2 * cta + 1) * BS * DS, cols + BS * DS, BLK_BYTES);
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
            bulk_store_1d(g0, cols, BLK_BYTES);
            bulk_store_1d(g0 + (size_t)BS * DS, cols + BS * DS, BLK_BYTES);
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
    const int dS = 64 * DK + 4;  // PLCfg<DK>::DS
    PsdLargeArgs a;

    a.d_pad = (int)ru(dmax, 2 * bs);
    a.CT = a.d_pad / (2 * bs);
    a.work_stride = ru((int64_t)(PL_NBUF + 1) * a.d_pad * dS + a.d_pad, 16);  // G (PL_NBUF version buffers), M, lambda
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
    static const bool want_prof = getenv("FOS_PSD_PROF") != nullptr;
    static unsigned long long *d_prof = nullptr;
    if (want_prof && !d_prof) FOS_CUDA(cudaMalloc((void **)&d_prof, 8 * sizeof(unsigned long long)));
    if (want_prof) FOS_CUDA(cudaMemsetAsync(d_prof, 0, 8 * sizeof(unsigned long long), h->stream));
    a.prof = want_prof ? d_prof : nullptr;
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
        // G buffers 1..PL_NBUF-1 must not hold anything a reader could take for their first version (bit 0): all ones
        FOS_CUDA(cudaMemset2DAsync(K.psd_work.p + (size_t)a.d_pad * dS, (size_t)a.work_stride * sizeof(double), 0xFF,
                                   (size_t)(PL_NBUF - 1) * a.d_pad * dS * sizeof(double), (size_t)n, h->stream));
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
    if (want_prof) {
        unsigned long long hp[8];
        FOS_CUDA(cudaStreamSynchronize(h->stream));
        FOS_CUDA(cudaMemcpy(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost));
        const double n = (double)std::max<unsigned long long>(hp[5], 1);
        fprintf(stderr, "psd_large d<=%d: %llu CTA-steps, cycles per step: load %.0f gram %.0f rotations %.0f apply %.0f sync %.0f\n",
                64 * DK, hp[5], hp[0] / n, hp[1] / n, hp[2] / n, hp[3] / n, hp[4] / n);
    }
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
