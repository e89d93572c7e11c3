// matop.cu -- device storage of A and the launch logic of the fused dual mat-vec (K1).
#include <dlfcn.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include "solver.cuh"

namespace fos {

// ---------------------------------------------------------------------------------------
// TMA descriptor through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn()
{
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FOS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p)
        throw Error(FOS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    fn = (PFN_encodeTiled)p;
    return fn;
}

static void make_tmap(CUtensorMap *map, const double *A, int64_t rows, int64_t cols, int64_t lda)
{
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)lda * 8};
    cuuint32_t box[2] = {(cuuint32_t)K1_BOXC, (cuuint32_t)K1_TR};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)A, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw Error(FOS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
}

static size_t k1_smem_bytes(int NV)
{
    return (size_t)K1_STAGES * K1_TILE_BYTES + (size_t)K1_STAGES * NV * K1_TR * 8 +
           (size_t)2 * K1_CONSUMER_WARPS * NV * K1_TR * 8 + 2 * K1_STAGES * 8 + 64;
}

void MatOp::init_dense(int64_t m_, int64_t n_, const double *Asrc, int64_t lda_src, int location, int64_t row_begin_,
                       int64_t row_count, int grid_ctas, cudaStream_t st)
{
    (void)st;
    kind = 1;
    m = m_;
    n = n_;
    n_pad = ru(n, PAD);
    m_pad = ru(m, PAD);
    row_begin = row_begin_;
    m_local = row_count;
    m_pad_local = ru(std::max<int64_t>(m_local, 1), PAD);
    FOS_REQUIRE(row_begin >= 0 && row_begin + m_local <= m, "row shard out of range");
    FOS_REQUIRE(row_begin % PAD == 0, "row_begin must be a multiple of 16");
    FOS_REQUIRE(lda_src >= n, "lda < n");
    const bool adoptable =
        location == FOS_MEM_DEVICE && (lda_src % 2 == 0) && ((reinterpret_cast<uintptr_t>(Asrc) & 15) == 0);
    if (adoptable) {
        A = Asrc;  // borrowed: the caller keeps the tensor alive for the lifetime of the problem
        lda = lda_src;
    } else {
        lda = ru(n, PAD);
        A_own.alloc((size_t)std::max<int64_t>(m_local, 1) * lda, false);
        FOS_CUDA(cudaMemset(A_own.p, 0, (size_t)std::max<int64_t>(m_local, 1) * lda * 8));
        FOS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));  // see DevBuf::alloc
        if (m_local > 0)
            FOS_CUDA(cudaMemcpy2D(A_own.p, (size_t)lda * 8, Asrc, (size_t)lda_src * 8, (size_t)n * 8, (size_t)m_local,
                                  location == FOS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
        FOS_SYNC_LEGACY();
        A = A_own.p;
    }
    make_tmap(&tmap, A, std::max<int64_t>(m_local, 1), n, lda);
    int G = grid_ctas > 0 ? grid_ctas : num_sms;
    by_smid = false;
    sm_weight.clear();
    plan = k1_make_plan(std::max<int64_t>(m_local, 1), n, G);
    upload_plan();
    rowpart.alloc((size_t)plan.NB * 2 * m_pad_local);
    // plain path / exchange buffers
    full_ax.alloc((size_t)2 * m_pad);
    full_atw.alloc((size_t)2 * n_pad);
    nchunk = (int)((std::max<int64_t>(m_local, 1) + K1S_ROWCHUNK - 1) / K1S_ROWCHUNK);
    scratch.alloc((size_t)nchunk * 2 * n_pad);
    d_one_band.upload(std::vector<int32_t>{0, 1});
    FOS_CUDA(cudaFuncSetAttribute(k1_dual_matvec_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)k1_smem_bytes(1)));
    FOS_CUDA(cudaFuncSetAttribute(k1_dual_matvec_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)k1_smem_bytes(2)));
    if (nranks > 1) xbuf.alloc((size_t)2 * (n_pad + m_pad));
    if (k1_balance && impl == 0 && grid_ctas <= 0) calibrate(st);
}

void MatOp::upload_plan()
{
    d_unit_begin.upload(plan.cta_unit_begin);
    d_slot_base.upload(plan.band_slot_base);
    d_first_cta.upload(plan.band_first_cta);
    colpart.alloc((size_t)std::max(plan.nslots, 1) * 2 * K1_BW);
    d_claim.alloc((size_t)plan.G + 8);
    d_cta_cycles.alloc((size_t)plan.G);
    d_cta_smid.alloc((size_t)plan.G);
}

// ---------------------------------------------------------------------------------------
// SM-speed calibration (option "k1_balance", off by default).  With a persistent CTA per SM and an even split of the
// tiles, ncu shows the SMs finishing between 0.85 and 0.99 of the kernel's duration.  This code binds the work ranges to
// SMs (K1Args::by_smid) and sizes them to the SMs' measured speed: a few timed passes over the real matrix at load,
// w_s ~ tiles_s / cycles_s, re-plan, repeat; the weights are cached per (device, shard shape) for the life of the
// process, so every handle of that shape gets the same plan and bit-identical results.
// MEASURED RESULT (profiles/r2_k1_balance.md): the spread of the per-SM times drops from 13 % to 10 % (full C2 matrix)
// and from 18 % to 12 % (2500-row shard), the pass time does not move (0.9383 -> 0.9384 ms, 0.1259 -> 0.1254 ms): an SM
// that finishes early simply leaves its share of the HBM bandwidth to the others.  The pass is bound by HBM (6.9 TB/s,
// 84 % of the DRAM peak ncu reports), not by the slowest SM -- which is why this stays opt-in.
// ---------------------------------------------------------------------------------------
namespace {
std::mutex g_calib_mutex;
std::map<std::tuple<int, int64_t, int64_t, int>, std::vector<double>> g_calib_cache;
}  // namespace

void MatOp::calibrate(cudaStream_t st)
{
    const int G = plan.G;
    const int64_t units = (int64_t)plan.NB * plan.RT;
    if ((kind != 1 && kind != 3) || G != num_sms || units < (k1_balance >= 2 ? 1 : 4) * (int64_t)G) return;  // too little work per SM
    int dev = 0;
    FOS_CUDA(cudaGetDevice(&dev));
    const auto key = std::make_tuple(dev, m_local, n, G);
    {
        std::lock_guard<std::mutex> lk(g_calib_mutex);
        auto it = g_calib_cache.find(key);
        if (it != g_calib_cache.end()) {
            sm_weight = it->second;
            if (!sm_weight.empty()) {
                plan = k1_make_plan(m_local, n, G, sm_weight.data());
                upload_plan();
                by_smid = true;
            }
            return;
        }
    }
    // zero vectors: the timing does not depend on the values
    DevBuf<double> zx, zw;
    zx.alloc((size_t)n_pad);
    // run() expects W indexed by GLOBAL row: the hybrid layout (kind 3) reads all of it, a shard only its own rows
    zw.alloc((size_t)(kind == 3 ? m_pad : m_pad_local) + K1_TR);
    const double *wbase = kind == 3 ? zw.p : zw.p - row_begin;
    const double *X[2] = {zx.p, zx.p}, *W[2] = {wbase, wbase};
    std::vector<double> w((size_t)G, 1.0);
    std::vector<unsigned long long> cyc((size_t)G);
    std::vector<int32_t> smid((size_t)G);
    bool ok = true;
    const bool keep_profile = profile;
    profile = false;
    Stats *keep_stats = stats;
    stats = nullptr;
    for (int round = 0; round < 4 && ok; round++) {
        by_smid = true;
        double tsum[2] = {0.0, 0.0};
        std::vector<double> t((size_t)G, 0.0);
        for (int rep = 0; rep < 3; rep++) {  // first repetition warms up
            run(2, X, W, nullptr, st);
            FOS_CUDA(cudaStreamSynchronize(st));
            FOS_CUDA(cudaMemcpy(cyc.data(), d_cta_cycles.p, (size_t)G * 8, cudaMemcpyDeviceToHost));
            FOS_CUDA(cudaMemcpy(smid.data(), d_cta_smid.p, (size_t)G * 4, cudaMemcpyDeviceToHost));
            for (int g = 0; g < G; g++)
                if (smid[(size_t)g] != g) ok = false;  // SM ids are not 0..G-1 one to one: keep the even split
            if (rep > 0)
                for (int g = 0; g < G; g++) t[(size_t)g] += (double)cyc[(size_t)g];
        }
        if (!ok) break;
        double tmean = 0.0, tmin = 1e300, tmax = 0.0;
        for (int g = 0; g < G; g++) {
            tmean += t[(size_t)g] / G;
            tmin = std::min(tmin, t[(size_t)g]);
            tmax = std::max(tmax, t[(size_t)g]);
        }
        (void)tsum;
        const double spread = (tmax - tmin) / tmean;
        if (round == 0) calib_spread[0] = spread;
        calib_spread[1] = spread;
        if (round == 3) break;
        // work_g ~ w_g and time_g = work_g / speed_g: move w_g towards equal times, damped
        double wmean = 0.0;
        for (int g = 0; g < G; g++) {
            w[(size_t)g] *= std::pow(tmean / t[(size_t)g], 0.7);
            wmean += w[(size_t)g] / G;
        }
        for (int g = 0; g < G; g++) w[(size_t)g] = std::min(1.3, std::max(0.7, w[(size_t)g] / wmean));
        plan = k1_make_plan(m_local, n, G, w.data());
        upload_plan();
    }
    profile = keep_profile;
    stats = keep_stats;
    if (!ok) {
        by_smid = false;
        w.clear();
        plan = k1_make_plan(m_local, n, G);
        upload_plan();
    }
    sm_weight = w;
    std::lock_guard<std::mutex> lk(g_calib_mutex);
    g_calib_cache[key] = w;
}

void MatOp::init_sparse(int64_t m_, int64_t n_, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                        int64_t base, cudaStream_t st)
{
    (void)st;
    kind = 2;
    m = m_;
    n = n_;
    n_pad = ru(n, PAD);
    m_pad = ru(m, PAD);
    row_begin = 0;
    m_local = m;
    m_pad_local = m_pad;
    FOS_REQUIRE(nranks == 1, "row sharding is implemented for dense storage only");
    nnz = colptr[n] - base;
    FOS_REQUIRE(nnz >= 0 && nnz < (int64_t)2147483647, "nnz out of int32 range");
    std::vector<int32_t> cptr((size_t)n + 1), cidx((size_t)nnz);
    std::vector<double> cval((size_t)nnz);
    for (int64_t j = 0; j <= n; j++) cptr[(size_t)j] = (int32_t)(colptr[j] - base);
    std::vector<int32_t> rcount((size_t)m + 1, 0);
    for (int64_t k = 0; k < nnz; k++) {
        int64_t i = rowval[k] - base;
        FOS_REQUIRE(i >= 0 && i < m, "row index out of range in CSC input");
        cidx[(size_t)k] = (int32_t)i;
        cval[(size_t)k] = nzval[k];
        rcount[(size_t)i + 1]++;
    }
    std::vector<int32_t> rptr((size_t)m + 1, 0);
    for (int64_t i = 0; i < m; i++) rptr[(size_t)i + 1] = rptr[(size_t)i] + rcount[(size_t)i + 1];
    std::vector<int32_t> fill(rptr.begin(), rptr.end() - 1), ridx((size_t)nnz);
    std::vector<double> rval((size_t)nnz);
    for (int64_t j = 0; j < n; j++)
        for (int32_t k = cptr[(size_t)j]; k < cptr[(size_t)j + 1]; k++) {
            int32_t i = cidx[(size_t)k];
            int32_t pos = fill[(size_t)i]++;
            ridx[(size_t)pos] = (int32_t)j;
            rval[(size_t)pos] = cval[(size_t)k];
        }
    csc_ptr.upload(cptr);
    csc_idx.upload(cidx);
    csc_val.upload(cval);
    csr_ptr.upload(rptr);
    csr_idx.upload(ridx);
    csr_val.upload(rval);
    full_ax.alloc((size_t)2 * m_pad);
    full_atw.alloc((size_t)2 * n_pad);
    d_one_band.upload(std::vector<int32_t>{0, 1});
}

// ---------------------------------------------------------------------------------------
// hybrid row storage: dense row block + CSR/CSC remainder (single problem, single rank)
// ---------------------------------------------------------------------------------------
namespace {

// non-zeros per row: one warp per row
__global__ void __launch_bounds__(256) k_rows_nnz(const double *__restrict__ A, int64_t lda, int64_t m, int64_t n,
                                                  int32_t *__restrict__ nnz)
{
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    const int lane = threadIdx.x & 31;
    const double *r = A + (size_t)row * lda;
    int c = 0;
    for (int64_t j = lane; j < n; j += 32) c += r[j] != 0.0 ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) nnz[row] = c;
}

// CSR of the listed rows: one warp per row, non-zeros compacted in column order (ballot prefix)
__global__ void __launch_bounds__(32) k_rows_extract(const double *__restrict__ A, int64_t lda, int64_t n,
                                                     const int32_t *__restrict__ row_id, const int32_t *__restrict__ row_ptr,
                                                     int32_t *__restrict__ col, double *__restrict__ val)
{
    const int sr = blockIdx.x;
    const int32_t p0 = row_ptr[sr];
    const int lane = threadIdx.x;
    const double *r = A + (size_t)row_id[sr] * lda;
    int32_t pos = p0;
    for (int64_t j0 = 0; j0 < n; j0 += 32) {
        const int64_t j = j0 + lane;
        const double v = j < n ? r[j] : 0.0;
        const unsigned int msk = __ballot_sync(0xffffffffu, v != 0.0);
        if (v != 0.0) {
            const int32_t q = pos + __popc(msk & ((1u << lane) - 1u));
            col[q] = (int32_t)j;
            val[q] = v;
        }
        pos += __popc(msk);
    }
}

// complete results of a hybrid pass: xbuf = [A' W (NV x n_pad) | A X (NV x m_pad)]
//   A' W = fold of the dense block's column partials + the CSC part;  A X = fold of the dense block's row
//   partials inside the block, the CSR part outside
template <int NV>
__global__ void __launch_bounds__(VBLOCK)
k_hybrid_combine(MVView V, int64_t n, int64_t n_pad, int64_t m_local, int64_t row_begin, int64_t m_pad,
                 const double *__restrict__ sp_ax, const double *__restrict__ sp_atw, double *__restrict__ xbuf,
                 const int32_t *skip_flag)
{
    if (skip_flag != nullptr && *skip_flag != 0) return;
    const int64_t total = n_pad + m_pad;
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < total; e += (int64_t)gridDim.x * VBLOCK) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
            if (e < n_pad) {
                xbuf[(size_t)v * n_pad + e] = e < n ? add_(mv_atw(V, v, e), sp_atw[(size_t)v * n_pad + e]) : 0.0;
            } else {
                const int64_t row = e - n_pad, lr = row - row_begin;
                xbuf[(size_t)NV * n_pad + (size_t)v * m_pad + row] =
                    (lr >= 0 && lr < m_local) ? mv_ax(V, v, lr) : sp_ax[(size_t)v * m_pad + row];
            }
        }
    }
}

}  // namespace

// Host-only decision of the hybrid row storage (unit-tested on CPU through fos_hybrid_plan): the dense block is
// [r0, r0 + md), r0 a multiple of 16 (the W slice K1 reads must stay 128-byte aligned), and holds every row with
// more than n/8 non-zeros; rows outside it with at least one entry go to CSR + CSC.
HybridPlan hybrid_row_plan(const int32_t *row_nnz, int64_t m, int64_t n)
{
    HybridPlan hp{};
    if (m < 2 * PAD || n < 1) return hp;
    const int64_t thr = std::max<int64_t>(1, n / 8);
    int64_t first = -1, last = -1;
    for (int64_t i = 0; i < m; i++)
        if (row_nnz[i] > thr) {
            if (first < 0) first = i;
            last = i;
        }
    if (first < 0) return hp;  // nothing dense: the caller should have chosen sparse storage
    hp.r0 = (first / PAD) * PAD;
    hp.md = last + 1 - hp.r0;
    for (int64_t i = 0; i < m; i++) {
        if ((i >= hp.r0 && i < hp.r0 + hp.md) || row_nnz[i] == 0) continue;
        hp.sparse_rows++;
        hp.sparse_nnz += row_nnz[i];
    }
    if (hp.sparse_nnz >= (int64_t)2147483647) return hp;
    const double dense_bytes = 8.0 * (double)m * (double)n;
    const double hybrid_bytes = 8.0 * (double)hp.md * (double)n + 24.0 * (double)hp.sparse_nnz;
    hp.use = hybrid_bytes <= 0.98 * dense_bytes;  // otherwise not worth three more launches per pass
    return hp;
}

void MatOp::init_hybrid(int64_t m_, int64_t n_, const double *Asrc, int64_t lda_src, int location, int grid_ctas,
                        cudaStream_t st, bool auto_only)
{
    FOS_REQUIRE(nranks == 1, "hybrid row storage is not offered with row sharding");
    init_dense(m_, n_, Asrc, lda_src, location, 0, m_, grid_ctas, st);  // A / lda: the whole matrix on the device
    hyb_sparse_rows = 0;
    if (impl != 0 || m < 2 * PAD) return;
    // 1. classify the rows
    DevBuf<int32_t> d_nnz;
    d_nnz.alloc((size_t)m);
    k_rows_nnz<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(A, lda, m, n, d_nnz.p);
    FOS_CUDA(cudaStreamSynchronize(st));
    std::vector<int32_t> rn((size_t)m);
    FOS_CUDA(cudaMemcpy(rn.data(), d_nnz.p, rn.size() * 4, cudaMemcpyDeviceToHost));
    HybridPlan hp = hybrid_row_plan(rn.data(), m, n);
    if (!hp.use) return;  // nothing to gain: the matrix stays dense (kind 1)
    if (auto_only) {
        const double saved = 8.0 * (double)m * (double)n - (8.0 * (double)hp.md * (double)n + 24.0 * (double)hp.sparse_nnz);
        if (saved < 64.0 * 1024.0 * 1024.0) return;  // three more launches per pass would cost more than the bytes saved
    }
    const int64_t r0 = hp.r0, r1 = hp.r0 + hp.md, md = hp.md, nz = hp.sparse_nnz;
    std::vector<int32_t> row_id, row_ptr(1, 0);
    {
        int64_t acc = 0;
        for (int64_t i = 0; i < m; i++) {
            if ((i >= r0 && i < r1) || rn[(size_t)i] == 0) continue;
            row_id.push_back((int32_t)i);
            acc += rn[(size_t)i];
            row_ptr.push_back((int32_t)acc);
        }
    }
    // 2. CSR of the remainder on the device, CSC on the host (nz entries)
    const int64_t NS = (int64_t)row_id.size();
    std::vector<int32_t> h_col((size_t)nz);
    std::vector<double> h_val((size_t)nz);
    if (nz > 0) {
        DevBuf<int32_t> d_id, d_ptr, d_col;
        DevBuf<double> d_val;
        d_id.upload(row_id);
        d_ptr.upload(row_ptr);
        d_col.alloc((size_t)nz, false);
        d_val.alloc((size_t)nz, false);
        k_rows_extract<<<(unsigned)NS, 32, 0, st>>>(A, lda, n, d_id.p, d_ptr.p, d_col.p, d_val.p);
        FOS_CUDA(cudaStreamSynchronize(st));
        FOS_CUDA(cudaMemcpy(h_col.data(), d_col.p, h_col.size() * 4, cudaMemcpyDeviceToHost));
        FOS_CUDA(cudaMemcpy(h_val.data(), d_val.p, h_val.size() * 8, cudaMemcpyDeviceToHost));
    }
    // full-size pointer arrays (rows of the dense block and empty rows have no entries): spmv_rows indexes by row
    std::vector<int32_t> rptr((size_t)m + 1, 0), cptr((size_t)n + 1, 0);
    for (int64_t s = 0; s < NS; s++) rptr[(size_t)row_id[(size_t)s] + 1] = row_ptr[(size_t)s + 1] - row_ptr[(size_t)s];
    for (int64_t i = 0; i < m; i++) rptr[(size_t)i + 1] += rptr[(size_t)i];
    for (int64_t q = 0; q < nz; q++) cptr[(size_t)h_col[(size_t)q] + 1]++;
    for (int64_t j = 0; j < n; j++) cptr[(size_t)j + 1] += cptr[(size_t)j];
    std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1), cidx((size_t)nz);
    std::vector<double> cval((size_t)nz);
    for (int64_t s = 0; s < NS; s++)  // rows in increasing order: every column's entries end up sorted by row
        for (int32_t q = row_ptr[(size_t)s]; q < row_ptr[(size_t)s + 1]; q++) {
            const int32_t pos = fill[(size_t)h_col[(size_t)q]]++;
            cidx[(size_t)pos] = row_id[(size_t)s];
            cval[(size_t)pos] = h_val[(size_t)q];
        }
    csr_ptr.upload(rptr);
    csr_idx.upload(h_col);  // the compacted rows are stored in row order: same layout as a CSR over all m rows
    csr_val.upload(h_val);
    csc_ptr.upload(cptr);
    csc_idx.upload(cidx);
    csc_val.upload(cval);
    nnz = nz;
    hyb_sparse_rows = NS;
    // 3. K1 works on the dense block only
    row_begin = r0;
    m_local = md;
    m_pad_local = ru(md, PAD);
    A = A + (size_t)r0 * lda;  // A_own (if any) keeps the ownership of the whole allocation
    make_tmap(&tmap, A, md, n, lda);
    const int G = grid_ctas > 0 ? grid_ctas : num_sms;
    by_smid = false;   // the block K1 streams has changed: even split (the weights were measured on the whole matrix)
    sm_weight.clear();
    plan = k1_make_plan(md, n, G);
    upload_plan();
    rowpart.alloc((size_t)plan.NB * 2 * m_pad_local);
    xbuf.alloc((size_t)2 * (n_pad + m_pad));
    kind = 3;
    if (k1_balance && grid_ctas <= 0) calibrate(st);
}

MatOp::~MatOp()
{
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
    p2p_close();
}

// ---------------------------------------------------------------------------------------
// fused exchange over peer memory: region = header | receive area [2 parities][nranks sources][2 vectors][n_pad+m_pad]
// ---------------------------------------------------------------------------------------
static int64_t p2p_recv_off() { return ru((int64_t)sizeof(P2PHeader), 256); }

void MatOp::p2p_export(uint8_t *handle_out)
{
    FOS_REQUIRE(kind == 1 && nranks > 1, "the peer-memory exchange needs a row-sharded dense problem (fos_comm_init + "
                                         "fos_load_conic_dense)");
    FOS_REQUIRE(nranks <= P2P_MAX_RANKS, "too many ranks for the peer-memory exchange");
    static_assert(sizeof(cudaIpcMemHandle_t) == FOS_IPC_HANDLE_BYTES, "IPC handle size");
    p2p_close();
    const size_t recv_bytes = (size_t)2 * nranks * 2 * (size_t)(n_pad + m_pad) * 8;
    const size_t bytes = (size_t)p2p_recv_off() + recv_bytes;
    p2p_region.alloc(bytes, false);
    // every receive slot starts as the sentinel (all ones): "nothing has arrived"
    FOS_CUDA(cudaMemset(p2p_region.p, 0xFF, bytes));
    FOS_SYNC_LEGACY();
    P2PHeader hd;
    memset(&hd, 0, sizeof(hd));
    hd.row_begin = row_begin;
    hd.m_local = m_local;
    FOS_CUDA(cudaMemcpy(p2p_region.p, &hd, sizeof(hd), cudaMemcpyHostToDevice));
    FOS_SYNC_LEGACY();
    p2p_local.alloc(4);
    cudaIpcMemHandle_t hdl;
    FOS_CUDA(cudaIpcGetMemHandle(&hdl, p2p_region.p));
    memcpy(handle_out, &hdl, sizeof(hdl));
}

void MatOp::p2p_import(const uint8_t *handles)
{
    FOS_REQUIRE(p2p_region.p != nullptr, "fos_comm_p2p_export must precede fos_comm_p2p_import");
    const int keep_mode = p2p.tail_flag_mode;
    memset(&p2p, 0, sizeof(p2p));
    p2p.tail_flag_mode = keep_mode;
    p2p.nranks = nranks;
    p2p.rank = rank;
    p2p.recv_off = p2p_recv_off();
    p2p.E = n_pad + m_pad;
    p2p.epoch = p2p_local.p;
    p2p.tickets = p2p_local.p + 1;
    p2p.error = p2p_local.p + 3;
    for (int r = 0; r < nranks; r++) {
        void *base = nullptr;
        if (r == rank) {
            base = p2p_region.p;
        } else {
            cudaIpcMemHandle_t hdl;
            memcpy(&hdl, handles + (size_t)r * FOS_IPC_HANDLE_BYTES, sizeof(hdl));
            cudaError_t e = cudaIpcOpenMemHandle(&base, hdl, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                throw Error(FOS_ERR_COMM, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) +
                                              ") failed: " + cudaGetErrorString(e));
            p2p_opened.push_back(base);
        }
        p2p.peer[r] = reinterpret_cast<unsigned char *>(base);
        P2PHeader hd;
        FOS_CUDA(cudaMemcpy(&hd, base, sizeof(hd), cudaMemcpyDeviceToHost));
        p2p.row_begin[r] = hd.row_begin;
        p2p.m_local[r] = hd.m_local;
    }
    p2p_on = true;
}

void MatOp::p2p_close()
{
    for (void *p : p2p_opened) cudaIpcCloseMemHandle(p);
    p2p_opened.clear();
    p2p_on = false;
}
void MatOp::prof_begin(int NV, cudaStream_t st)
{
    if (!profile) return;
    if (ev_used + 2 > ev_pool.size()) {
        for (int k = 0; k < 64; k++) {
            cudaEvent_t e;
            FOS_CUDA(cudaEventCreate(&e));
            ev_pool.push_back(e);
        }
    }
    ev_nv.resize(ev_pool.size() / 2);
    ev_nv[ev_used / 2] = NV;
    FOS_CUDA(cudaEventRecord(ev_pool[ev_used], st));
}
void MatOp::prof_end(cudaStream_t st)
{
    if (!profile) return;
    FOS_CUDA(cudaEventRecord(ev_pool[ev_used + 1], st));
    ev_used += 2;
}
void MatOp::prof_collect()
{
    if (!profile) return;
    // launches faster than 25 TB/s did no work: they are the predicated no-ops of a CG batch
    const double min_ms = bytes_per_pass() / 25e12 * 1e3;
    for (size_t k = 0; k + 1 < ev_used; k += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev_pool[k], ev_pool[k + 1]) != cudaSuccess) continue;
        const int nv = ev_nv[k / 2];
        if (nv == 0) {  // fused CG tail: predicated no-ops take ~2 us
            if (ms > 0.004f) {
                prof_ms[0] += ms;
                prof_n[0]++;
            }
            continue;
        }
        if ((double)ms < min_ms) {
            prof_skipped++;
            continue;
        }
        prof_ms[nv] += ms;
        prof_n[nv]++;
    }
    ev_used = 0;
}
void MatOp::prof_reset()
{
    ev_used = 0;
    for (int k = 0; k < 3; k++) {
        prof_ms[k] = 0;
        prof_n[k] = 0;
    }
    prof_skipped = 0;
}

double MatOp::bytes_per_pass() const
{
    if (kind == 1) return 8.0 * (double)m_local * (double)n;
    if (kind == 3) return 8.0 * (double)m_local * (double)n + 2.0 * 12.0 * (double)nnz;
    return 2.0 * 12.0 * (double)nnz;  // CSR pass + CSC pass, 8-byte value + 4-byte index
}

MVView MatOp::view_full(int NV, const double *ax, const double *atw) const
{
    (void)NV;
    MVView V;
    V.rowpart = ax;
    V.nb = 1;
    V.rp_sb = 0;
    V.rp_sv = m_pad;
    V.colpart = atw;
    V.slot_base = d_one_band.p;
    V.bw_shift = 62;
    V.bw_mask = (int64_t)0x3fffffffffffffffLL;
    V.cp_ss = 0;
    V.cp_sv = n_pad;
    return V;
}

template <int NV>
MVView MatOp::run_t(const double *const *X, const double *const *W, const int32_t *skip, cudaStream_t st,
                    bool defer_exchange)
{
    K1Args<NV> a;
    for (int v = 0; v < NV; v++) {
        a.X[v] = X[v];
        a.W[v] = W[v] + row_begin;
    }
    a.skip_flag = skip;
    a.n_pad = n_pad;
    a.m_pad_local = m_pad_local;
    MVView V;
    if ((kind == 1 && impl == 0) || kind == 3) {
        a.rowpart = rowpart.p;
        a.colpart = colpart.p;
        a.cta_unit_begin = d_unit_begin.p;
        a.band_slot_base = d_slot_base.p;
        a.band_first_cta = d_first_cta.p;
        a.RT = plan.RT;
        a.NB = plan.NB;
        a.kc_last = plan.kc_last;
        a.by_smid = by_smid ? 1 : 0;
        a.claim = d_claim.p;
        a.claim_epoch = d_claim.p + plan.G;
        a.exit_ticket = d_claim.p + plan.G + 1;
        a.cta_cycles = d_cta_cycles.p;
        a.cta_smid = d_cta_smid.p;
        k1_dual_matvec_tma<NV><<<plan.G, K1_THREADS, k1_smem_bytes(NV), st>>>(tmap, a);
        if (stats) stats->launches++;
        V.rowpart = rowpart.p;
        V.nb = plan.NB;
        V.rp_sb = (int64_t)NV * m_pad_local;
        V.rp_sv = m_pad_local;
        V.colpart = colpart.p;
        V.slot_base = d_slot_base.p;
        V.bw_shift = K1_BW_SHIFT;
        V.bw_mask = K1_BW - 1;
        V.cp_ss = (int64_t)NV * K1_BW;
        V.cp_sv = K1_BW;
    } else if (kind == 1) {
        // plain path: complete results for the local rows
        a.rowpart = full_ax.p;
        a.colpart = full_atw.p;
        a.m_pad_local = m_pad_local;
        const int wpb = 8;
        k1_simple_rows<NV><<<(unsigned)((m_local + wpb - 1) / wpb), 256, 0, st>>>(A, lda, m_local, n, a);
        dim3 g((unsigned)((n + 255) / 256), (unsigned)nchunk);
        k1_simple_cols<NV><<<g, 256, 0, st>>>(A, lda, m_local, n, a, scratch.p);
        k1_simple_cols_sum<NV><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, scratch.p, nchunk, n, n_pad);
        if (stats) stats->launches += 3;
        V = view_full(NV, full_ax.p, full_atw.p);
        V.rp_sv = m_pad_local;
    } else {
        a.rowpart = full_ax.p;
        a.colpart = full_atw.p;
        const int wpb = 8;
        spmv_rows<NV><<<(unsigned)((m + wpb - 1) / wpb), 256, 0, st>>>(csr_ptr.p, csr_idx.p, csr_val.p, m, a, 0, m_pad);
        spmv_rows<NV><<<(unsigned)((n + wpb - 1) / wpb), 256, 0, st>>>(csc_ptr.p, csc_idx.p, csc_val.p, n, a, 1, n_pad);
        if (stats) stats->launches += 2;
        V = view_full(NV, full_ax.p, full_atw.p);
    }
    if (kind == 3) {
        // the CSR / CSC remainder, then one fold that merges it with the dense block's partials
        K1Args<NV> sa = a;
        for (int v = 0; v < NV; v++) sa.W[v] = W[v];  // global rows
        sa.rowpart = full_ax.p;
        sa.colpart = full_atw.p;
        const int wpb = 8;
        spmv_rows<NV><<<(unsigned)((m + wpb - 1) / wpb), 256, 0, st>>>(csr_ptr.p, csr_idx.p, csr_val.p, m, sa, 0, m_pad);
        spmv_rows<NV><<<(unsigned)((n + wpb - 1) / wpb), 256, 0, st>>>(csc_ptr.p, csc_idx.p, csc_val.p, n, sa, 1, n_pad);
        const int64_t total = n_pad + m_pad;
        const int grid = (int)std::min<int64_t>((total + VBLOCK - 1) / VBLOCK, 4 * (int64_t)num_sms);
        k_hybrid_combine<NV><<<grid, VBLOCK, 0, st>>>(V, n, n_pad, m_local, row_begin, m_pad, full_ax.p, full_atw.p, xbuf.p,
                                                      skip);
        if (stats) stats->launches += 3;
        V = view_full(NV, xbuf.p + (size_t)NV * n_pad, xbuf.p);
    }
    if (stats && skip == nullptr) stats->total_passes++;  // predicated CG launches are counted by cg_solve
    if (nranks > 1 && p2p_on && defer_exchange) {
        // the consumer (k_cg_tail_hsde<true>) folds, publishes and gathers itself
    } else if (nranks > 1 && p2p_on) {
        // ONE kernel: fold the local partials, publish, gather from the peers over NVLink (k1_exchange_p2p).
        // Its blocks wait for each other and for the peers, so the grid must be co-resident.
        const int64_t total = n_pad + m_pad;
        int grid = (int)std::min<int64_t>((total + VBLOCK - 1) / VBLOCK, 2 * (int64_t)num_sms);
        k1_exchange_p2p<NV><<<grid, VBLOCK, 0, st>>>(V, p2p, n, n_pad, m_pad, xbuf.p, skip);
        if (stats) stats->launches++;
        V = view_full(NV, xbuf.p + (size_t)NV * n_pad, xbuf.p);
    } else if (nranks > 1) {
        // fold the local partials into the exchange buffer, all-reduce over NVLink, hand out
        // a complete view.  Rows owned by other ranks are zero in the local contribution.
        const int64_t total = n_pad + m_pad;
        int grid = (int)std::min<int64_t>((total + VBLOCK - 1) / VBLOCK, 4 * (int64_t)num_sms);
        k1_finalize_local<NV><<<grid, VBLOCK, 0, st>>>(V, n, n_pad, m_local, row_begin, m_pad, xbuf.p, skip);
        if (stats) stats->launches++;
        const size_t count = (size_t)NV * (size_t)(n_pad + m_pad);
        if (comm == nullptr)
            throw Error(FOS_ERR_COMM, "no NCCL communicator on this handle (fos_comm_init was given an all-zero id): "
                                      "enable the peer-memory exchange (fos_comm_p2p_export / _import) before running");
        int rc = nccl_api().AllReduce(xbuf.p, xbuf.p, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, st);
        if (rc != 0) throw Error(FOS_ERR_COMM, std::string("ncclAllReduce failed: ") + nccl_api().GetErrorString(rc));
        V = view_full(NV, xbuf.p + (size_t)NV * n_pad, xbuf.p);
    }
    return V;
}

MVView MatOp::run(int NV, const double *const *X, const double *const *W, const int32_t *skip, cudaStream_t st,
                  bool defer_exchange)
{
    FOS_REQUIRE(kind != 0, "matrix not loaded");
    prof_begin(NV, st);
    MVView V = (NV == 1) ? run_t<1>(X, W, skip, st, defer_exchange) : run_t<2>(X, W, skip, st, defer_exchange);
    prof_end(st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) throw Error(FOS_ERR_CUDA, std::string("mat-vec launch failed: ") + cudaGetErrorString(e));
    return V;
}

// ---------------------------------------------------------------------------------------
// NCCL via dlopen
// ---------------------------------------------------------------------------------------
NcclApi &nccl_api()
{
    static NcclApi api;
    if (api.lib) return api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names) {
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) throw Error(FOS_ERR_COMM, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
    auto sym = [&](const char *s) {
        void *p = dlsym(lib, s);
        if (!p) throw Error(FOS_ERR_COMM, std::string("NCCL symbol missing: ") + s);
        return p;
    };
    api.GetUniqueId = (int (*)(NcclId *))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, NcclId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))sym("ncclAllReduce");
    api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    api.lib = lib;
    return api;
}

}  // namespace fos
