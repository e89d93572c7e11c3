// matvec.cuh -- K1: the fused dual mat-vec.  One pass over A (dense row-major FP64, HBM
// resident) produces BOTH  A*X  (X: n x NV)  and  A'*W  (W: m x NV), NV = 1 or 2.
//
// Replaces the four serial CSC passes of one KKT product in the reference:
//   KKTMatrix.mul!   utilities/affinepluslinear.jl:45,47
//   -> HSDEMatrixQ.mul!  problemforms/HSDE/HSDEAffine.jl:51-52  (called twice)
//
// Design (B200, HBM-bound, no tensor cores -- 1 flop/byte):
//   * A is cut into tiles of TR x TC = 16 x 512 doubles (64 KB); KC = 4 tiles side by side
//     form a "row group" (16 x 2048).  Column bands of BW = 2048 columns are swept top to
//     bottom, band after band; the linear sequence of row groups is split evenly (by tile
//     count) over G persistent CTAs, one per SM.
//   * one producer warp streams tiles with TMA (cp.async.bulk.tensor.2d, two 16x256 boxes per
//     tile, zero fill at the edges) into a 3-stage shared-memory ring guarded by mbarriers;
//     the matching 16 entries of each W vector ride on the same barrier (cp.async.bulk).
//   * 8 consumer warps: warp w / lane l owns columns 64w+2l, +1 of every tile (one 128-bit
//     shared load per row, conflict free).  Column sums stay in registers for a whole band
//     sweep (no cross-thread reduction at all); row sums stay in registers across the KC tiles
//     of a row group, then one transposed warp-shuffle butterfly + an 8-way shared-memory
//     add produce the 16 x NV row results of that group.
//   * results are written once, as partials: rowpart[band][v][row], colpart[slot][v][col]
//     (slot = (band, k-th CTA touching it)).  The consumer kernel (K2) adds the <= ~20 partials
//     per entry in fixed order -> bitwise reproducible, no atomics.
//   * extra HBM traffic over the algorithmic 8*m*n bytes: 8*NV*(m*NB + n*slots/NB...) ~ 0.2 %.
#pragma once
#include <cmath>

#include "common.cuh"

namespace fos {

constexpr int K1_TR = 16;             // tile rows
constexpr int K1_BOXC = 256;          // TMA box width (hardware limit: 256 elements per dim)
constexpr int K1_TC = 512;            // tile columns = 2 boxes
constexpr int K1_KC = 4;              // tiles per row group
constexpr int K1_BW = K1_TC * K1_KC;  // band width = 2048 columns
constexpr int K1_BW_SHIFT = 11;
constexpr int K1_STAGES = 3;
constexpr int K1_CONSUMER_WARPS = 8;
constexpr int K1_THREADS = (K1_CONSUMER_WARPS + 1) * 32;
constexpr int K1_TILE_BYTES = K1_TR * K1_TC * 8;  // 65536

struct K1Plan {
    int32_t G = 0;        // CTAs
    int32_t RT = 0;       // row tiles (of the local rows)
    int32_t NB = 0;       // column bands
    int32_t nslots = 0;   // column-partial slots
    int32_t kc_last = 0;  // tiles per row group in the last band
    std::vector<int32_t> cta_unit_begin;  // [G+1], unit = band*RT + rowtile
    std::vector<int32_t> band_slot_base;  // [NB+1]
    std::vector<int32_t> band_first_cta;  // [NB]
};

// `weight` (optional, G_req entries, mean 1): relative speed of the SM that runs work range g (MatOp::calibrate):
// range g receives a share of the tiles proportional to its weight instead of 1/G.
inline K1Plan k1_make_plan(int64_t m_local, int64_t n, int G_req, const double *weight = nullptr)
{
    K1Plan P;
    P.RT = (int32_t)((m_local + K1_TR - 1) / K1_TR);
    P.NB = (int32_t)((n + K1_BW - 1) / K1_BW);
    int64_t last_cols = n - (int64_t)(P.NB - 1) * K1_BW;
    P.kc_last = (int32_t)((last_cols + K1_TC - 1) / K1_TC);
    int64_t units = (int64_t)P.NB * P.RT;
    int G = G_req;
    if (units < G) G = (int)units;
    if (G < 1) G = 1;
    P.G = G;
    int64_t Wt = (int64_t)P.RT * K1_KC * (P.NB - 1) + (int64_t)P.RT * P.kc_last;
    P.cta_unit_begin.assign(G + 1, 0);
    // unit u starts at cumulative weight pre(u); CTA g owns units with pre(u) in [g*Wt/G, (g+1)*Wt/G)
    auto pre = [&](int64_t u) -> int64_t {
        int64_t b = u / P.RT, rt = u % P.RT;
        int64_t kc = (b == P.NB - 1) ? P.kc_last : K1_KC;
        return (int64_t)P.RT * K1_KC * b + rt * kc;
    };
    // every CTA gets a non-empty, contiguous range (G <= units): the slot of a CTA inside a band is
    // its distance from the band's first CTA, which needs the owners of a band to be consecutive.
    int64_t u = 0;
    double wsum = 0.0, wacc = 0.0;
    if (weight && G == G_req)
        for (int g = 0; g < G; g++) wsum += weight[g];
    for (int g = 0; g < G; g++) {
        int64_t target = (Wt * g + G - 1) / G;  // ceil
        if (wsum > 0.0) {
            target = (int64_t)std::ceil((double)Wt * (wacc / wsum) - 1e-9);
            wacc += weight[g];
        }
        while (u < units - (G - g) && pre(u) < target) u++;
        if (g > 0 && u <= P.cta_unit_begin[g - 1]) u = P.cta_unit_begin[g - 1] + 1;
        P.cta_unit_begin[g] = (int32_t)u;
    }
    P.cta_unit_begin[G] = (int32_t)units;
    P.band_slot_base.assign(P.NB + 1, 0);
    P.band_first_cta.assign(P.NB, 0);
    int slot = 0;
    for (int b = 0; b < P.NB; b++) {
        int64_t ub = (int64_t)b * P.RT, ue = ub + P.RT;
        int first = -1, cnt = 0;
        for (int g = 0; g < G; g++) {
            int64_t s = P.cta_unit_begin[g], e = P.cta_unit_begin[g + 1];
            if (s < ue && e > ub && e > s) {
                if (first < 0) first = g;
                cnt++;
            }
        }
        P.band_first_cta[b] = first < 0 ? 0 : first;
        P.band_slot_base[b] = slot;
        slot += cnt;
    }
    P.band_slot_base[P.NB] = slot;
    P.nslots = slot;
    return P;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA (bulk tensor / bulk linear copies)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int32_t c0, int32_t c1,
                                            uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        :
        : "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// The same copy for data that is read once per pass (the batch kernel's A tiles), with the evict-first policy K1 uses for
// its tiles: the stream must not push out what IS reused (the problems' iteration vectors; ncu: DRAM traffic was 1.3x
// the tile bytes).
__device__ __forceinline__ void bulk_load_1d_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :
                 : "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void consumer_bar_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(K1_CONSUMER_WARPS * 32) : "memory");
}

template <int NV>
struct K1Args {
    const double *X[NV];  // n_pad entries each, 16-byte aligned, zero beyond n
    const double *W[NV];  // local rows, readable for RT*16 entries, 16-byte aligned
    double *rowpart;      // [NB][NV][m_pad_local]
    double *colpart;      // [nslots][NV][BW]
    const int32_t *cta_unit_begin;
    const int32_t *band_slot_base;
    const int32_t *band_first_cta;
    const int32_t *skip_flag;  // device int: nonzero -> return immediately (CG batch predication)
    // Work ranges bound to SMs (MatOp::calibrate): with by_smid the CTA running on SM s takes range s (one CTA per SM:
    // the kernel needs the whole shared memory), so that a range can be sized to the measured speed of its SM.
    // claim[r] == epoch + 1 marks range r as taken in this launch (a CTA whose SM's range is taken -- which only happens
    // when the grid is not fully co-resident -- probes for the next free one: every range runs exactly once whatever
    // the placement, and its content never depends on who runs it).
    int32_t by_smid;
    unsigned int *claim;        // [G]
    unsigned int *claim_epoch;  // launches completed so far; advanced by the last CTA to finish
    unsigned int *exit_ticket;
    unsigned long long *cta_cycles;  // optional [G]: SM cycles from kernel start to the end of range g's work
    int32_t *cta_smid;               // optional [G]: the SM that ran range g
    int64_t n_pad;
    int64_t m_pad_local;
    int32_t RT, NB, kc_last;
};

// Sum V = NV*TR per-lane values across the 32 lanes of a warp with a transposed butterfly:
// log2(V) halving exchange steps, then plain butterflies for the remaining lane bits.
// On return vals[0] of lane L holds the complete sum of value index (L >> (5 - log2 V)) when
// V <= 32.  Fixed order -> deterministic.
template <int V>
__device__ __forceinline__ void warp_transpose_reduce(double (&vals)[V], int lane)
{
    static_assert(V == 8 || V == 16 || V == 32, "V must be 8, 16 or 32");
    int s = 16;
#pragma unroll
    for (int cnt = V / 2; cnt >= 1; cnt >>= 1, s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < cnt; i++) {
            double send = upper ? vals[i] : vals[i + cnt];
            double keep = upper ? vals[i + cnt] : vals[i];
            vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
#pragma unroll
    for (; s >= 1; s >>= 1) vals[0] += __shfl_xor_sync(0xffffffffu, vals[0], s);
}

template <int NV>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_dual_matvec_tma(const __grid_constant__ CUtensorMap tmap, const K1Args<NV> args)
{
    if (args.skip_flag != nullptr && *args.skip_flag != 0) return;

    constexpr int TR = K1_TR, KC = K1_KC, S = K1_STAGES;
    constexpr int V = NV * TR;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double *tiles = reinterpret_cast<double *>(smem_raw);                           // S * 64 KB
    double *wbuf = reinterpret_cast<double *>(smem_raw + (size_t)S * K1_TILE_BYTES);  // S * NV * TR
    double *red = wbuf + S * NV * TR;                                               // 2 * 8 * V
    uint64_t *full = reinterpret_cast<uint64_t *>(red + 2 * K1_CONSUMER_WARPS * V);
    uint64_t *empty = full + S;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int s_rank;
    const long long t_start = clock64();
    unsigned int epoch = 0;
    if (threadIdx.x == 0) {
        int r = blockIdx.x;
        if (args.by_smid) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            epoch = *args.claim_epoch;
            r = smid < gridDim.x ? (int)smid : (int)blockIdx.x;
            for (unsigned int k = 0; k < gridDim.x; k++) {
                if (atomicExch(&args.claim[r], epoch + 1u) != epoch + 1u) break;  // range r was free: it is ours
                r = r + 1 == (int)gridDim.x ? 0 : r + 1;
            }
            if (args.cta_smid) args.cta_smid[r] = (int32_t)smid;
        }
        s_rank = r;
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], K1_CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int g = s_rank;
    const int u_begin = args.cta_unit_begin[g], u_end = args.cta_unit_begin[g + 1];
    // the CTA's last act (thread 0): record its time, and -- as the last CTA of the launch -- open the next epoch
    auto finish_cta = [&]() {
        if (args.cta_cycles) args.cta_cycles[g] = (unsigned long long)(clock64() - t_start);
        if (args.by_smid) {
            __threadfence();
            if (atomicAdd(args.exit_ticket, 1u) == gridDim.x - 1) {
                *args.exit_ticket = 0u;
                *args.claim_epoch = epoch + 1u;
                __threadfence();
            }
        }
    };
    if (u_begin >= u_end) {
        if (threadIdx.x == 0) finish_cta();
        return;
    }

    if (warp == K1_CONSUMER_WARPS) {
        // ================= producer warp: one elected lane drives TMA =================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            uint32_t t = 0;
            for (int u = u_begin; u < u_end; u++) {
                const int b = u / args.RT, rt = u - b * args.RT;
                const int kc = (b == args.NB - 1) ? args.kc_last : KC;
                for (int k = 0; k < kc; k++, t++) {
                    const int s = t % S;
                    const uint32_t ph = (t / S) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], K1_TILE_BYTES + NV * TR * 8);
                    double *dst = tiles + (size_t)s * (K1_TILE_BYTES / 8);
                    const int32_t col0 = b * K1_BW + k * K1_TC, row0 = rt * TR;
                    tma_load_2d(dst, &tmap, &full[s], col0, row0, pol);
                    tma_load_2d(dst + TR * K1_BOXC, &tmap, &full[s], col0 + K1_BOXC, row0, pol);
#pragma unroll
                    for (int v = 0; v < NV; v++)
                        bulk_load_1d(wbuf + (s * NV + v) * TR, args.W[v] + (int64_t)rt * TR, TR * 8, &full[s]);
                }
            }
        }
        return;
    }

    // ======================= consumer warps =======================
    const int boxsel = warp >> 2;                              // which 256-column box
    const int cin = ((warp & 3) << 6) + (lane << 1);           // column inside the box
    const int ccol = (boxsel << 8) + cin;                      // column inside the 512-wide tile
    const int toff = boxsel * (TR * K1_BOXC) + cin;            // offset of row 0 inside a stage

    double2 ca[KC][NV];
    double2 xv[KC][NV];
    double rs[V];
    int cur_band = -1;
    uint32_t t = 0;

    auto flush_cols = [&](int band) {
        const int slot = args.band_slot_base[band] + (g - args.band_first_cta[band]);
        double *base = args.colpart + (size_t)slot * (NV * K1_BW);
#pragma unroll
        for (int k = 0; k < KC; k++)
#pragma unroll
            for (int v = 0; v < NV; v++)
                *reinterpret_cast<double2 *>(base + v * K1_BW + k * K1_TC + ccol) = ca[k][v];
    };

    for (int u = u_begin; u < u_end; u++) {
        const int b = u / args.RT, rt = u - b * args.RT;
        const int kc = (b == args.NB - 1) ? args.kc_last : KC;
        if (b != cur_band) {
            if (cur_band >= 0) flush_cols(cur_band);
            cur_band = b;
#pragma unroll
            for (int k = 0; k < KC; k++) {
                const int64_t col = (int64_t)b * K1_BW + k * K1_TC + ccol;
#pragma unroll
                for (int v = 0; v < NV; v++) {
                    ca[k][v] = make_double2(0.0, 0.0);
                    xv[k][v] = (col < args.n_pad) ? *reinterpret_cast<const double2 *>(args.X[v] + col)
                                                  : make_double2(0.0, 0.0);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < V; i++) rs[i] = 0.0;

#pragma unroll
        for (int k = 0; k < KC; k++) {
            if (k < kc) {
                const int s = t % S;
                const uint32_t ph = (t / S) & 1;
                mbar_wait(&full[s], ph);
                const double *tile = tiles + (size_t)s * (K1_TILE_BYTES / 8) + toff;
                const double *sw = wbuf + s * NV * TR;
#pragma unroll
                for (int r = 0; r < TR; r++) {
                    const double2 a = *reinterpret_cast<const double2 *>(tile + r * K1_BOXC);
#pragma unroll
                    for (int v = 0; v < NV; v++) {
                        const double w = sw[v * TR + r];
                        rs[v * TR + r] = fma(a.x, xv[k][v].x, rs[v * TR + r]);
                        rs[v * TR + r] = fma(a.y, xv[k][v].y, rs[v * TR + r]);
                        ca[k][v].x = fma(a.x, w, ca[k][v].x);
                        ca[k][v].y = fma(a.y, w, ca[k][v].y);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                t++;
            }
        }

        // ---- row sums of this group: lanes -> warps -> global partial ----
        warp_transpose_reduce<V>(rs, lane);
        double *rbuf = red + (u & 1) * (K1_CONSUMER_WARPS * V);
        if (V == 32) rbuf[warp * V + lane] = rs[0];
        else if ((lane & 1) == 0) rbuf[warp * V + (lane >> 1)] = rs[0];
        consumer_bar_sync();
        if (threadIdx.x < V) {
            double sum = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < K1_CONSUMER_WARPS; w8++) sum += rbuf[w8 * V + threadIdx.x];
            const int v = threadIdx.x / TR, r = threadIdx.x - v * TR;
            args.rowpart[((size_t)b * NV + v) * args.m_pad_local + (size_t)rt * TR + r] = sum;
        }
    }
    if (cur_band >= 0) flush_cols(cur_band);
    consumer_bar_sync();                 // every consumer warp has finished its part of the range
    if (threadIdx.x == 0) finish_cta();  // (thread 0 is a consumer: the producer warp is the last one)
}

// ---------------------------------------------------------------------------------------
// plain two-kernel path ("matvec_impl" = 1): no TMA, no fusion; exists to validate K1 on the
// device and as the ncu comparison point.  Produces a complete (1 band / 1 slot) result.
// ---------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
k1_simple_rows(const double *__restrict__ A, int64_t lda, int64_t m_local, int64_t n, K1Args<NV> args)
{
    if (args.skip_flag != nullptr && *args.skip_flag != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= m_local) return;
    const double *a = A + row * lda;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
    for (int64_t j = lane; j < n; j += 32) {
        const double aij = a[j];
#pragma unroll
        for (int v = 0; v < NV; v++) acc[v] = fma(aij, args.X[v][j], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const double sres = warp_sum(acc[v]);
        if (lane == 0) args.rowpart[(size_t)v * args.m_pad_local + row] = sres;
    }
}

constexpr int K1S_ROWCHUNK = 1024;
template <int NV>
__global__ void __launch_bounds__(256)
k1_simple_cols(const double *__restrict__ A, int64_t lda, int64_t m_local, int64_t n, K1Args<NV> args,
               double *__restrict__ scratch /* [nchunk][NV][n_pad] */)
{
    if (args.skip_flag != nullptr && *args.skip_flag != 0) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int64_t r0 = (int64_t)blockIdx.y * K1S_ROWCHUNK;
    const int64_t r1 = r0 + K1S_ROWCHUNK < m_local ? r0 + K1S_ROWCHUNK : m_local;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
    for (int64_t i = r0; i < r1; i++) {
        const double aij = A[i * lda + j];
#pragma unroll
        for (int v = 0; v < NV; v++) acc[v] = fma(aij, args.W[v][i], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; v++) scratch[((size_t)blockIdx.y * NV + v) * args.n_pad + j] = acc[v];
}
template <int NV>
__global__ void __launch_bounds__(256)
k1_simple_cols_sum(K1Args<NV> args, const double *__restrict__ scratch, int nchunk, int64_t n, int64_t cp_sv)
{
    if (args.skip_flag != nullptr && *args.skip_flag != 0) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        double s = 0.0;
        for (int c = 0; c < nchunk; c++) s += scratch[((size_t)c * NV + v) * args.n_pad + j];
        args.colpart[(size_t)v * cp_sv + j] = s;
    }
}

// ---------------------------------------------------------------------------------------
// sparse path: CSR for A*X (one warp per row), CSC for A'*W (one warp per column); int32
// indices, 12 bytes per non-zero per direction.  Complete (1 band / 1 slot) result.
// ---------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
spmv_rows(const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx, const double *__restrict__ val,
          int64_t nrows, K1Args<NV> args, int which /*0: A*X, 1: A'*W*/, int64_t out_sv)
{
    if (args.skip_flag != nullptr && *args.skip_flag != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
    const int32_t p0 = ptr[row], p1 = ptr[row + 1];
    for (int32_t p = p0 + lane; p < p1; p += 32) {
        const double a = val[p];
        const int32_t j = idx[p];
#pragma unroll
        for (int v = 0; v < NV; v++) acc[v] = fma(a, (which == 0 ? args.X[v] : args.W[v])[j], acc[v]);
    }
    double *out = which == 0 ? args.rowpart : args.colpart;
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const double sres = warp_sum(acc[v]);
        if (lane == 0) out[(size_t)v * out_sv + row] = sres;
    }
}
#endif  // __CUDACC__

}  // namespace fos
