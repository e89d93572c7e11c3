// capi.cu -- the extern "C" boundary declared in include/fos_b200.h.  Exceptions never cross
// it: every entry point converts them to an int32 code and stores the text on the handle.
#include <algorithm>
#include <mutex>

#include "solver.cuh"

using namespace fos;

struct fos_handle_s {
    Handle h;
};

static std::string g_last_error;
static std::mutex g_err_mutex;

static int32_t fail(fos_handle_t hh, int32_t code, const std::string &msg)
{
    if (hh) hh->h.err = msg;
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_last_error = msg;
    return code;
}

#define FOS_API_BEGIN(hh)                                                  \
    if (!(hh)) return fail(nullptr, FOS_ERR_INVALID, "null handle");       \
    try {                                                                  \
        cudaSetDevice((hh)->h.device);

#define FOS_API_END(hh)                                                    \
        return FOS_OK;                                                     \
    } catch (const Error &e) {                                             \
        cudaGetLastError();                                                \
        return fail((hh), e.code, e.what());                               \
    } catch (const std::bad_alloc &) {                                     \
        return fail((hh), FOS_ERR_NOMEM, "host allocation failed");        \
    } catch (const std::exception &e) {                                    \
        return fail((hh), FOS_ERR_INVALID, e.what());                      \
    }

extern "C" {

int32_t fos_abi_version(void) { return FOS_ABI_VERSION; }

int32_t fos_create(fos_handle_t *out, int32_t device)
{
    if (!out) return fail(nullptr, FOS_ERR_INVALID, "null output pointer");
    *out = nullptr;
    fos_handle_s *hh = nullptr;
    try {
        hh = new fos_handle_s();
        hh->h.create(device);
        *out = hh;
        return FOS_OK;
    } catch (const Error &e) {
        int32_t rc = fail(nullptr, e.code, e.what());
        delete hh;
        return rc;
    } catch (const std::exception &e) {
        int32_t rc = fail(nullptr, FOS_ERR_INVALID, e.what());
        delete hh;
        return rc;
    }
}

int32_t fos_destroy(fos_handle_t hh)
{
    if (!hh) return FOS_OK;
    delete hh;
    return FOS_OK;
}

const char *fos_last_error(fos_handle_t hh)
{
    if (hh) return hh->h.err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mutex);
    return g_last_error.c_str();
}

int32_t fos_set_option(fos_handle_t hh, const char *key, double value)
{
    FOS_API_BEGIN(hh)
    FOS_REQUIRE(key != nullptr, "null key");
    Handle &h = hh->h;
    const std::string k(key);
    h.drop_graphs();  // most options change kernel arguments or the kernel sequence of the captured iteration
    if (k == "matvec_impl") {
        FOS_REQUIRE(value == 0 || value == 1, "matvec_impl must be 0 or 1");
        h.matvec_impl = (int)value;
        h.A.impl = (int)value;
    } else if (k == "grid_ctas") {
        FOS_REQUIRE(!h.loaded, "grid_ctas must be set before loading the problem");
        h.grid_ctas = (int)value;
    } else if (k == "cg_batch") {
        h.cg_batch = (int)value;
    } else if (k == "fuse_rhs") {
        h.fuse_rhs = value != 0;
    } else if (k == "tail_flags") {
        FOS_REQUIRE(value == 0 || value == 1, "tail_flags must be 0 (block to block) or 1 (per rank)");
        h.A.p2p.tail_flag_mode = (int)value;
        h.tail_flag_mode = (int)value;
    } else if (k == "tail_blocks") {
        FOS_REQUIRE(value >= 0 && value <= 2 * h.num_sms, "tail_blocks out of range");
        h.tail_blocks = (int)value;
    } else if (k == "psd_warm") {
        h.cones.psd_warm_enabled = value != 0;
    } else if (k == "fuse_tail") {
        h.fuse_tail = value != 0;
    } else if (k == "profile_matvec") {
        h.A.profile = value != 0;
        h.A.prof_reset();
    } else if (k == "exchange_impl") {
        FOS_REQUIRE(value == 0 || value == 1, "exchange_impl must be 0 (NCCL) or 1 (peer memory)");
        FOS_REQUIRE(value == 0 || h.A.p2p.nranks > 1, "exchange_impl = 1 needs fos_comm_p2p_import first");
        h.A.p2p_on = value != 0;
    } else if (k == "batch_hybrid") {
        FOS_REQUIRE(!(h.batch && h.batch->loaded), "batch_hybrid must be set before loading the batch");
        h.batch_hybrid = value != 0;
    } else if (k == "hybrid_rows") {
        FOS_REQUIRE(!h.loaded, "hybrid_rows must be set before loading the problem");
        FOS_REQUIRE(value == 0 || value == 1 || value == 2, "hybrid_rows must be 0 (off), 1 (whenever it saves bytes) or 2 (auto)");
        h.hybrid_rows = (int)value;
    } else if (k == "batch_ctas") {
        h.batch_ctas = (int)value;
        if (h.batch) h.batch->grid_ctas = (int)value;
    } else if (k == "k1_balance") {
        FOS_REQUIRE(!h.loaded, "k1_balance must be set before loading the problem");
        FOS_REQUIRE(value == 0 || value == 1 || value == 2, "k1_balance must be 0 (even split), 1 (auto) or 2 (always)");
        h.A.k1_balance = (int)value;
    } else if (k == "psd_warp_max_d") {
        FOS_REQUIRE(value >= 0 && value <= PSD_WARP_MAX_D, "psd_warp_max_d must be 0 .. 16");
        g_psd_warp_max_d = (int)value;  // applies to problems loaded afterwards
    } else if (k == "tail_trace") {
        if (value != 0) {
            h.tail_trace.alloc(48);
            h.gbar.trace = h.tail_trace.p;
        } else {
            h.gbar.trace = nullptr;
        }
    } else if (k == "use_graphs") {
        h.use_graphs = value != 0;
    } else {
        throw Error(FOS_ERR_INVALID, "unknown option: " + k);
    }
    FOS_API_END(hh)
}

// ---- multi-GPU --------------------------------------------------------------------------
int32_t fos_comm_unique_id(uint8_t *id_out)
{
    if (!id_out) return fail(nullptr, FOS_ERR_INVALID, "null id buffer");
    try {
        NcclId id;
        memset(&id, 0, sizeof(id));
        int rc = nccl_api().GetUniqueId(&id);
        if (rc != 0) throw Error(FOS_ERR_COMM, std::string("ncclGetUniqueId: ") + nccl_api().GetErrorString(rc));
        memcpy(id_out, id.bytes, FOS_COMM_ID_BYTES);
        return FOS_OK;
    } catch (const Error &e) {
        return fail(nullptr, e.code, e.what());
    }
}

int32_t fos_comm_init(fos_handle_t hh, int32_t rank, int32_t nranks, const uint8_t *idb)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    FOS_REQUIRE(!h.loaded, "fos_comm_init must precede loading");
    FOS_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    FOS_REQUIRE(idb != nullptr, "null id");
    bool want_nccl = false;  // an all-zero id = "peer-memory exchange only": no NCCL communicator is created
    for (int k = 0; k < FOS_COMM_ID_BYTES; k++) want_nccl = want_nccl || idb[k] != 0;
    if (nranks > 1 && want_nccl) {
        NcclId id;
        memcpy(id.bytes, idb, FOS_COMM_ID_BYTES);
        void *comm = nullptr;
        int rc = nccl_api().CommInitRank(&comm, nranks, id, rank);
        if (rc != 0) throw Error(FOS_ERR_COMM, std::string("ncclCommInitRank: ") + nccl_api().GetErrorString(rc));
        h.comm = comm;
        h.A.comm = comm;
    }
    h.rank = rank;
    h.nranks = nranks;
    h.A.rank = rank;
    h.A.nranks = nranks;
    FOS_API_END(hh)
}

int32_t fos_comm_p2p_export(fos_handle_t hh, uint8_t *handle_out)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(handle_out != nullptr, "null handle buffer");
    FOS_CUDA(cudaStreamSynchronize(h.stream));
    h.A.p2p_export(handle_out);
    FOS_API_END(hh)
}

int32_t fos_comm_p2p_import(fos_handle_t hh, const uint8_t *handles)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(handles != nullptr, "null handle table");
    h.drop_graphs();
    h.A.p2p_import(handles);
    FOS_API_END(hh)
}

// ---- loading ----------------------------------------------------------------------------
static bool want_dense(int32_t storage, int64_t m, int64_t n, int64_t nnz)
{
    if (storage == FOS_STORAGE_DENSE) return true;
    if (storage == FOS_STORAGE_SPARSE) return false;
    return (double)nnz > 0.25 * (double)m * (double)n;
}

static void load_matrix_csc(Handle &h, int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval,
                            const double *nzval, int64_t base, int32_t storage)
{
    FOS_REQUIRE(m >= 0 && n >= 0 && (m > 0 || n > 0), "empty problem");
    FOS_REQUIRE(colptr && (rowval || colptr[n] == base) && (nzval || colptr[n] == base), "null CSC arrays");
    FOS_REQUIRE(base == 0 || base == 1, "index_base must be 0 or 1");
    const int64_t nnz = colptr[n] - base;
    h.A.impl = h.matvec_impl;
    if (want_dense(storage, m, n, nnz) && m > 0 && n > 0) {
        FOS_REQUIRE(h.nranks == 1, "row sharding needs fos_load_conic_dense");
        std::vector<double> D((size_t)m * (size_t)n, 0.0);
        for (int64_t j = 0; j < n; j++)
            for (int64_t k = colptr[j] - base; k < colptr[j + 1] - base; k++) {
                const int64_t i = rowval[k] - base;
                FOS_REQUIRE(i >= 0 && i < m, "row index out of range in CSC input");
                D[(size_t)i * (size_t)n + (size_t)j] += nzval[k];
            }
        if (h.hybrid_rows)
            h.A.init_hybrid(m, n, D.data(), n, FOS_MEM_HOST, h.grid_ctas, h.stream, h.hybrid_rows == 2);
        else
            h.A.init_dense(m, n, D.data(), n, FOS_MEM_HOST, 0, m, h.grid_ctas, h.stream);
    } else {
        h.A.init_sparse(m, n, colptr, rowval, nzval, base, h.stream);
    }
}

int32_t fos_load_conic_csc(fos_handle_t hh, int64_t m, int64_t n, const int64_t *colptr, const int64_t *rowval,
                           const double *nzval, int64_t index_base, const double *b, const double *c,
                           int64_t ncones1, const int32_t *cone_type1, const int64_t *cone_len1, int64_t ncones2,
                           const int32_t *cone_type2, const int64_t *cone_len2, int32_t storage)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.loaded = false;
    FOS_REQUIRE((b || m == 0) && (c || n == 0), "null b / c");
    load_matrix_csc(h, m, n, colptr, rowval, nzval, index_base, storage);
    h.load_conic(m, n, b, c, ncones1, cone_type1, cone_len1, ncones2, cone_type2, cone_len2);
    h.begin_solve();
    FOS_API_END(hh)
}

int32_t fos_load_conic_dense(fos_handle_t hh, int64_t m, int64_t n, const double *A, int64_t lda, int32_t a_location,
                             int64_t row_begin, int64_t row_count, const double *b, const double *c,
                             int64_t ncones1, const int32_t *cone_type1, const int64_t *cone_len1, int64_t ncones2,
                             const int32_t *cone_type2, const int64_t *cone_len2)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.loaded = false;
    FOS_REQUIRE(m > 0 && n > 0 && A && b && c, "bad dense problem arguments");
    FOS_REQUIRE(a_location == FOS_MEM_HOST || a_location == FOS_MEM_DEVICE, "bad a_location");
    // a device-resident matrix was produced on the caller's stream(s); the library works on its own
    // non-blocking stream, so wait for everything the caller has enqueued before reading it
    if (a_location == FOS_MEM_DEVICE) FOS_CUDA(cudaDeviceSynchronize());
    h.A.impl = h.matvec_impl;
    if (h.hybrid_rows && h.nranks == 1 && row_begin == 0 && row_count == m)
        h.A.init_hybrid(m, n, A, lda, a_location, h.grid_ctas, h.stream, h.hybrid_rows == 2);
    else
        h.A.init_dense(m, n, A, lda, a_location, row_begin, row_count, h.grid_ctas, h.stream);
    h.load_conic(m, n, b, c, ncones1, cone_type1, cone_len1, ncones2, cone_type2, cone_len2);
    h.begin_solve();
    FOS_API_END(hh)
}

int32_t fos_load_affine_csc(fos_handle_t hh, int64_t am, int64_t an, const int64_t *colptr, const int64_t *rowval,
                            const double *nzval, int64_t index_base, const double *b, const double *q, int32_t beta,
                            int32_t decreasing_accuracy, int64_t ncones, const int32_t *cone_type,
                            const int64_t *cone_len, int32_t storage)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.loaded = false;
    load_matrix_csc(h, am, an, colptr, rowval, nzval, index_base, storage);
    h.load_affine(am, an, b, q, beta, decreasing_accuracy, ncones, cone_type, cone_len);
    h.begin_solve();
    FOS_API_END(hh)
}

// ---- algorithm and iterate -----------------------------------------------------------------
int32_t fos_set_algorithm(fos_handle_t hh, int32_t alg, double alpha, double alpha1, double alpha2, double beta,
                          int64_t iproj)
{
    FOS_API_BEGIN(hh)
    FOS_REQUIRE(hh->h.loaded || (hh->h.batch && hh->h.batch->loaded), "no problem loaded on this handle");
    hh->h.set_algorithm(alg, alpha, alpha1, alpha2, beta, iproj);
    FOS_API_END(hh)
}

int32_t fos_set_box(fos_handle_t hh, int64_t start, int64_t len, double lo, double hi)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(h.L.form == 1, "IndBox applies to the S2 of the Feasibility form");
    FOS_REQUIRE(start >= 0 && len >= 0 && start + len <= h.N, "box range out of bounds");
    // logical [x(an); z(am)] -> padded offsets; a range may straddle the x / z boundary
    const int64_t an = h.L.n;
    const int64_t a0 = start, a1 = start + len;
    if (a0 < an) h.cones.set_box(a0, std::min(a1, an) - a0, lo, hi);
    if (a1 > an) h.cones.set_box(h.L.n_pad + std::max<int64_t>(a0 - an, 0), a1 - std::max(a0, an), lo, hi);
    FOS_API_END(hh)
}

int32_t fos_set_linesearch(fos_handle_t hh, int64_t lsinterval)
{
    FOS_API_BEGIN(hh)
    hh->h.require_loaded();
    FOS_REQUIRE(lsinterval >= 0, "lsinterval must be >= 0");
    hh->h.lsinterval = lsinterval;
    FOS_API_END(hh)
}

int32_t fos_set_direct(fos_handle_t hh, int32_t on)
{
    FOS_API_BEGIN(hh)
    hh->h.set_direct(on != 0);
    FOS_API_END(hh)
}

int64_t fos_iterate_length(fos_handle_t hh)
{
    if (!hh) return -1;
    if (hh->h.batch && hh->h.batch->loaded) return hh->h.batch->N;
    if (!hh->h.loaded) return -1;
    return hh->h.N;
}

int32_t fos_set_iterate(fos_handle_t hh, const double *z, int64_t len)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(z && len == h.N, "iterate length mismatch");
    h.pack_from_host(z, h.x.p);
    FOS_API_END(hh)
}

int32_t fos_set_initial_iterate(fos_handle_t hh)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    std::vector<double> z((size_t)h.N, 0.0);
    if (h.L.form == 0) {  // HSDE_getinitialvalue (HSDE.jl:40-47)
        z[(size_t)(h.N / 2 - 1)] = 1.0;
        z[(size_t)(h.N - 1)] = 1.0;
    }
    h.pack_from_host(z.data(), h.x.p);
    FOS_API_END(hh)
}

int32_t fos_get_iterate(fos_handle_t hh, double *z, int64_t len)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(z && len == h.N, "iterate length mismatch");
    h.unpack_to_host(h.x.p, z);
    FOS_API_END(hh)
}

int32_t fos_get_state(fos_handle_t hh, int32_t which, double *buf, int64_t len)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(buf && len == h.N, "state length mismatch");
    const double *src = nullptr;
    switch (which) {
    case 0: src = h.x.p; break;
    case 1: src = h.tmp1.p; break;
    case 2: src = h.tmp2.p; break;
    case 3: src = h.sol.p; break;
    case 4: src = h.rhs.p; break;
    case 5: src = h.proj.p; break;
    case 6: src = h.fy.p; break;
    case 7: src = h.dp.p; break;
    case 8: src = h.dq.p; break;
    default: throw Error(FOS_ERR_INVALID, "unknown state selector");
    }
    h.unpack_to_host(src, buf);
    FOS_API_END(hh)
}

int32_t fos_set_state(fos_handle_t hh, int32_t which, const double *buf, int64_t len)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(buf && len == h.N, "state length mismatch");
    double *dst = nullptr;
    switch (which) {
    case 0: dst = h.x.p; break;
    case 3:
        dst = h.sol.p;
        h.firstrun = false;
        break;
    case 6: dst = h.fy.p; break;
    case 7: dst = h.dp.p; break;
    case 8: dst = h.dq.p; break;
    default: throw Error(FOS_ERR_INVALID, "this state vector cannot be set");
    }
    h.pack_from_host(buf, dst);
    FOS_API_END(hh)
}

int32_t fos_set_info(fos_handle_t hh, int32_t which, double value)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    switch (which) {
    case 0: h.s1_calls = (int64_t)value; break;
    case 2:
        h.sync_ctrl();
        h.h_ctrl->alpha12 = value;
        FOS_CUDA(cudaMemcpyAsync(h.d_ctrl.p, h.h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, h.stream));
        FOS_CUDA(cudaStreamSynchronize(h.stream));
        break;
    case 3: h.fista_t = value; break;
    default: throw Error(FOS_ERR_INVALID, "this scalar cannot be set");
    }
    FOS_API_END(hh)
}

int32_t fos_get_info(fos_handle_t hh, int32_t which, double *out)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    FOS_REQUIRE(out != nullptr, "null output");
    switch (which) {
    case 0: *out = (double)h.s1_calls; break;
    case 1: *out = (double)h.cgiter; break;
    case 2:
        h.sync_ctrl();
        *out = h.h_ctrl->alpha12;
        break;
    case 3: *out = h.fista_t; break;
    case 4: *out = h.warn_maxit ? 1.0 : 0.0; break;
    case 5: *out = (double)h.stats.total_cg; break;
    case 6: *out = (double)h.stats.total_passes; break;
    case 7: *out = (double)h.stats.launches; break;
    case 8:
        h.sync_ctrl();
        *out = h.h_ctrl->ls_alphabest;
        break;
    case 9: *out = h.A.prof_ms[2]; break;
    case 10: *out = (double)h.A.prof_n[2]; break;
    case 11: *out = h.A.prof_ms[1]; break;
    case 12: *out = (double)h.A.prof_n[1]; break;
    case 13: *out = (double)h.A.prof_skipped; break;
    case 14: *out = (h.batch && h.batch->loaded) ? h.batch->bytes_per_pass() : h.A.bytes_per_pass(); break;
    case 15: *out = (double)h.num_sms; break;
    case 16: *out = h.A.prof_ms[0]; break;
    case 17: *out = (double)h.A.prof_n[0]; break;
    case 18: *out = (double)h.A.kind; break;
    case 19: *out = h.A.kind == 3 ? (double)h.A.m_local : 0.0; break;
    case 20: *out = h.A.kind == 3 ? (double)h.A.hyb_sparse_rows : 0.0; break;
    case 21: *out = h.A.by_smid ? 1.0 : 0.0; break;
    case 22: *out = h.A.calib_spread[0]; break;
    case 23: *out = h.A.calib_spread[1]; break;
    default: throw Error(FOS_ERR_INVALID, "unknown info selector");
    }
    FOS_API_END(hh)
}

int32_t fos_get_tail_trace(fos_handle_t hh, double *out)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    FOS_REQUIRE(out != nullptr, "null output");
    FOS_REQUIRE(h.gbar.trace != nullptr, "set the option \"tail_trace\" first");
    FOS_CUDA(cudaStreamSynchronize(h.stream));
    unsigned long long t[48];
    FOS_CUDA(cudaMemcpy(t, h.tail_trace.p, sizeof(t), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 48; k++) out[k] = (double)t[k];
    FOS_API_END(hh)
}

// ---- the hot loop --------------------------------------------------------------------------
int32_t fos_begin_solve(fos_handle_t hh)
{
    FOS_API_BEGIN(hh)
    hh->h.require_loaded();
    hh->h.begin_solve();
    FOS_API_END(hh)
}

int32_t fos_run(fos_handle_t hh, int64_t i_start, int64_t n_iters, int64_t checki, double eps, int64_t *iters_done,
                int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec, double *trace)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    int64_t done = h.run(i_start, n_iters, checki, eps, records, records ? rec_cap : 0, n_rec, trace);
    if (iters_done) *iters_done = done;
    if (status) *status = h.status;
    FOS_API_END(hh)
}

int32_t fos_finish(fos_handle_t hh, double *guess, int64_t len, double *record, int64_t *n_rec, int32_t *status)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(!guess || len == h.N, "guess length mismatch");
    h.finish(guess, record, n_rec);
    if (status) *status = h.status;
    FOS_API_END(hh)
}

int32_t fos_solve(fos_handle_t hh, int64_t max_iters, int64_t checki, double eps, double *guess, int64_t len,
                  int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(!guess || len == h.N, "guess length mismatch");
    h.begin_solve();
    int64_t nr1 = 0, nr2 = 0;
    int64_t done = h.run(1, max_iters, checki, eps, records, records ? rec_cap : 0, &nr1, nullptr);
    double *tail = (records && nr1 < rec_cap) ? records + nr1 * FOS_REC_LEN : nullptr;
    h.finish(guess, tail, &nr2);
    if (iters_done) *iters_done = done;
    if (n_rec) *n_rec = nr1 + nr2;
    if (status) *status = h.status == FOS_STATUS_CONTINUE ? FOS_STATUS_INDETERMINATE : h.status;
    FOS_API_END(hh)
}

// ---- batch mode -----------------------------------------------------------------------------
static BatchSolver &batch_of(Handle &h)
{
    FOS_REQUIRE(h.batch && h.batch->loaded, "the handle is not in batch mode (fos_load_conic_dense_batch)");
    return *h.batch;
}
static int batch_vec_of(int32_t which, bool for_set)
{
    switch (which) {
    case 0: return BV_X;
    case 1: FOS_REQUIRE(!for_set, "this state vector cannot be set"); return BV_TMP1;
    case 2: FOS_REQUIRE(!for_set, "this state vector cannot be set"); return BV_TMP2;
    case 3: return BV_SOL;
    case 5: FOS_REQUIRE(!for_set, "this state vector cannot be set"); return BV_PROJ;
    case 6: return BV_FY;
    case 7: return BV_DP;
    case 8: return BV_DQ;
    default: throw Error(FOS_ERR_INVALID, "unknown state selector");
    }
}

int32_t fos_load_conic_dense_batch(fos_handle_t hh, int64_t nprob, int64_t m, int64_t n, const double *A, int64_t lda,
                                   int64_t pstride, int32_t a_location, const double *b, const double *c,
                                   int64_t ncones1, const int32_t *cone_type1, const int64_t *cone_len1,
                                   int64_t ncones2, const int32_t *cone_type2, const int64_t *cone_len2)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    FOS_REQUIRE(A && b && c, "null problem data");
    FOS_REQUIRE(lda >= n && pstride >= (m - 1) * lda + n, "bad leading dimension / problem stride");
    FOS_REQUIRE(a_location == FOS_MEM_HOST || a_location == FOS_MEM_DEVICE, "bad a_location");
    FOS_REQUIRE(h.nranks == 1, "batch mode is split across GPUs by the caller (one handle per rank), not row-sharded");
    if (a_location == FOS_MEM_DEVICE) FOS_CUDA(cudaDeviceSynchronize());  // see fos_load_conic_dense
    h.loaded = false;
    h.batch.reset(new BatchSolver());
    h.batch->grid_ctas = h.batch_ctas;
    h.batch->hybrid = h.batch_hybrid != 0;
    h.batch->load(&h, nprob, m, n, A, lda, pstride, a_location, b, c, ncones1, cone_type1, cone_len1, ncones2,
                  cone_type2, cone_len2);
    h.alg = FOS_ALG_GAP;
    h.alpha = 0.8;
    h.alpha1 = 1.8;
    h.alpha2 = 1.8;
    h.betaA = 0.0;
    FOS_API_END(hh)
}

int64_t fos_batch_size(fos_handle_t hh)
{
    if (!hh || !hh->h.batch || !hh->h.batch->loaded) return -1;
    return hh->h.batch->B;
}

int32_t fos_set_iterate_batch(fos_handle_t hh, const double *z)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    if (z) bs.set_vector(BV_X, z, 0, bs.B);
    else bs.set_initial_iterate();
    FOS_API_END(hh)
}

int32_t fos_get_iterate_batch(fos_handle_t hh, double *z)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    FOS_REQUIRE(z != nullptr, "null output");
    bs.get_vector(BV_X, z, 0, bs.B);
    FOS_API_END(hh)
}

int32_t fos_get_state_batch(fos_handle_t hh, int32_t which, double *buf)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    FOS_REQUIRE(buf != nullptr, "null output");
    bs.get_vector(batch_vec_of(which, false), buf, 0, bs.B);
    FOS_API_END(hh)
}

int32_t fos_set_state_batch(fos_handle_t hh, int32_t which, const double *buf)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    FOS_REQUIRE(buf != nullptr, "null input");
    bs.set_vector(batch_vec_of(which, true), buf, 0, bs.B);
    if (which == 3) {  // restoring the CG warm start clears S1's first-run flag (affinepluslinear.jl:101-104)
        FOS_CUDA(cudaStreamSynchronize(hh->h.stream));
        std::vector<BatchCtl> hc((size_t)bs.B);
        FOS_CUDA(cudaMemcpy(hc.data(), bs.dctl.p, hc.size() * sizeof(BatchCtl), cudaMemcpyDeviceToHost));
        for (BatchCtl &c : hc) c.firstrun = 0;
        FOS_CUDA(cudaMemcpy(bs.dctl.p, hc.data(), hc.size() * sizeof(BatchCtl), cudaMemcpyHostToDevice));
        FOS_SYNC_LEGACY();
    }
    FOS_API_END(hh)
}

int32_t fos_get_info_batch(fos_handle_t hh, int32_t which, double *out)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    FOS_REQUIRE(out != nullptr, "null output");
    FOS_CUDA(cudaStreamSynchronize(hh->h.stream));
    std::vector<BatchCtl> hc((size_t)bs.B);
    FOS_CUDA(cudaMemcpy(hc.data(), bs.dctl.p, hc.size() * sizeof(BatchCtl), cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < bs.B; p++) {
        const BatchCtl &c = hc[(size_t)p];
        switch (which) {
        case 0: out[p] = (double)c.s1_calls; break;
        case 1: out[p] = (double)c.cgiter; break;
        case 2: out[p] = c.alpha12; break;
        case 3: out[p] = c.fista_t; break;
        case 4: out[p] = (double)c.warn_maxit; break;
        case 5: out[p] = (double)c.total_cg; break;
        case 6: out[p] = (double)c.total_passes; break;
        default: throw Error(FOS_ERR_INVALID, "unknown info selector");
        }
    }
    FOS_API_END(hh)
}

int32_t fos_set_info_batch(fos_handle_t hh, int32_t which, const double *values)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    FOS_REQUIRE(values != nullptr, "null input");
    FOS_CUDA(cudaStreamSynchronize(hh->h.stream));
    std::vector<BatchCtl> hc((size_t)bs.B);
    FOS_CUDA(cudaMemcpy(hc.data(), bs.dctl.p, hc.size() * sizeof(BatchCtl), cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < bs.B; p++) {
        BatchCtl &c = hc[(size_t)p];
        switch (which) {
        case 0: c.s1_calls = (int64_t)values[p]; break;
        case 2: c.alpha12 = values[p]; break;
        case 3: c.fista_t = values[p]; break;
        default: throw Error(FOS_ERR_INVALID, "this scalar cannot be set");
        }
    }
    FOS_CUDA(cudaMemcpy(bs.dctl.p, hc.data(), hc.size() * sizeof(BatchCtl), cudaMemcpyHostToDevice));
    FOS_SYNC_LEGACY();
    FOS_API_END(hh)
}

int32_t fos_begin_solve_batch(fos_handle_t hh)
{
    FOS_API_BEGIN(hh)
    batch_of(hh->h).begin_solve();
    FOS_API_END(hh)
}

int32_t fos_run_batch(fos_handle_t hh, int64_t i_start, int64_t n_iters, int64_t checki, double eps,
                      int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    bs.launch(i_start, n_iters, checki, eps, true, false);
    bs.collect(iters_done, status, records, records ? rec_cap : 0, n_rec, nullptr);
    FOS_API_END(hh)
}

int32_t fos_finish_batch(fos_handle_t hh, double *guess, double *record, int64_t *n_rec, int32_t *status)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    bs.launch(0, 0, 1 << 30, bs.last_eps, false, true);
    bs.collect(nullptr, status, record, record ? 1 : 0, n_rec, nullptr);
    if (guess) bs.get_vector(BV_PROJ, guess, 0, bs.B);
    FOS_API_END(hh)
}

int32_t fos_solve_batch(fos_handle_t hh, int64_t max_iters, int64_t checki, double eps, double *guess,
                        int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap, int64_t *n_rec)
{
    FOS_API_BEGIN(hh)
    BatchSolver &bs = batch_of(hh->h);
    bs.begin_solve();
    bs.launch(1, max_iters, checki, eps, true, true);
    bs.collect(iters_done, status, records, records ? rec_cap : 0, n_rec, nullptr);
    if (status)
        for (int64_t p = 0; p < bs.B; p++)
            if (status[p] == FOS_STATUS_CONTINUE) status[p] = FOS_STATUS_INDETERMINATE;  // HSDE.jl:56-59
    if (guess) bs.get_vector(BV_PROJ, guess, 0, bs.B);
    FOS_API_END(hh)
}

// ---- unit-level entry points ------------------------------------------------------------------
int32_t fos_a_mul(fos_handle_t hh, const double *x, double *y, int32_t transpose)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(x && y, "null vector");
    const int64_t n = h.L.n, m = h.L.m, n_pad = h.L.n_pad, m_pad = h.L.m_pad;
    // stage X (n_pad) and W (m_pad) in w1; the unused side is zero
    FOS_CUDA(cudaMemsetAsync(h.w1.p, 0, (size_t)h.L.NP * 8, h.stream));
    double *dX = h.w1.p, *dW = h.w1.p + n_pad;
    FOS_REQUIRE(n_pad + m_pad <= h.L.NP, "internal: staging too small");
    FOS_CUDA(cudaMemcpyAsync(transpose ? dW : dX, x, (size_t)(transpose ? m : n) * 8, cudaMemcpyHostToDevice,
                             h.stream));
    const double *X[1] = {dX};
    const double *W[1] = {dW};
    MVView V = h.A.run(1, X, W, nullptr, h.stream);
    // reuse the plain KKT epilogue with a zero "in" to fold the partials: out1 = A'W, out2 = A X
    Lay Lp = h.L;
    Lp.form = 1;
    Lp.NP = n_pad + m_pad;
    FOS_CUDA(cudaMemsetAsync(h.w3.p, 0, (size_t)h.L.NP * 8, h.stream));
    FOS_LAUNCH(&h, k2_kkt_plain<K2_OUT>, h.vgrid(Lp.NP), VBLOCK, 0, Lp, V, h.w3.p, h.w2.p, nullptr, nullptr, nullptr,
               h.d_ctrl.p, h.rb, 0);
    FOS_CUDA(cudaMemcpyAsync(y, transpose ? h.w2.p : h.w2.p + n_pad, (size_t)(transpose ? n : m) * 8,
                             cudaMemcpyDeviceToHost, h.stream));
    FOS_CUDA(cudaStreamSynchronize(h.stream));
    FOS_API_END(hh)
}

int32_t fos_q_mul(fos_handle_t hh, const double *B, double *Y, int32_t transpose)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(h.L.form == 0 && B && Y, "fos_q_mul needs the conic form");
    // B has length l: pack it as the first half of an iterate
    std::vector<double> z((size_t)h.N, 0.0), out((size_t)h.N, 0.0);
    memcpy(z.data(), B, (size_t)(h.N / 2) * 8);
    h.pack_from_host(z.data(), h.w1.p);
    FOS_CUDA(cudaMemsetAsync(h.w2.p, 0, (size_t)h.L.NP * 8, h.stream));
    h.q_mul(h.w1.p, h.w2.p, transpose != 0);
    h.unpack_to_host(h.w2.p, out.data());
    memcpy(Y, out.data(), (size_t)(h.N / 2) * 8);
    FOS_API_END(hh)
}

int32_t fos_kkt_mul(fos_handle_t hh, const double *x, double *y)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(x && y, "null vector");
    h.pack_from_host(x, h.w1.p);
    FOS_CUDA(cudaMemsetAsync(h.w2.p, 0, (size_t)h.L.NP * 8, h.stream));
    h.kkt_mul(h.w1.p, h.w2.p);
    h.unpack_to_host(h.w2.p, y);
    FOS_API_END(hh)
}

int32_t fos_affine_prox(fos_handle_t hh, const double *x, double *y)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(x && y, "null vector");
    h.pack_from_host(x, h.w1.p);
    h.s1_prox(h.w1.p);
    h.sol_scaled_to(h.w2.p);
    h.unpack_to_host(h.w2.p, y);
    FOS_API_END(hh)
}

int32_t fos_hsdematrix_prox(fos_handle_t hh, const double *x, double *y)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(h.L.form == 0 && x && y, "fos_hsdematrix_prox needs the conic form");
    const size_t bytes = (size_t)h.L.NP * 8;
    // borrow S1's CG buffers; keep its warm start, rhs and counters intact
    FOS_CUDA(cudaMemcpyAsync(h.w3.p, h.sol.p, bytes, cudaMemcpyDeviceToDevice, h.stream));
    const int64_t keep_cgiter = h.cgiter;
    const int64_t keep_total = h.stats.total_cg;
    h.pack_from_host(x, h.rhs.p);                                                              // rhs = x
    FOS_CUDA(cudaMemcpyAsync(h.sol.p, h.rhs.p, bytes, cudaMemcpyDeviceToDevice, h.stream));   // :109-114 first run
    h.cgiter = 1;
    h.cg_solve((double)h.N * 2.220446049250313e-16, 1000);                                     // :106, :116
    FOS_CUDA(cudaMemcpyAsync(h.w2.p, h.sol.p, bytes, cudaMemcpyDeviceToDevice, h.stream));
    h.q_mul(h.sol.p, h.w2.p + h.L.LP, false);                                                  // :120-124  v = Q u
    h.unpack_to_host(h.w2.p, y);
    FOS_CUDA(cudaMemcpyAsync(h.sol.p, h.w3.p, bytes, cudaMemcpyDeviceToDevice, h.stream));
    FOS_CUDA(cudaMemsetAsync(h.rhs.p, 0, bytes, h.stream));  // HSDE: rhs2 = b = 0 (HSDE.jl:22); rhs1 is rebuilt per prox
    FOS_CUDA(cudaStreamSynchronize(h.stream));
    h.cgiter = keep_cgiter;
    h.stats.total_cg = keep_total;
    FOS_API_END(hh)
}

int32_t fos_cone_prox(fos_handle_t hh, const double *x, double *y)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE(x && y, "null vector");
    h.pack_from_host(x, h.w1.p);
    EpiArgs E{};
    h.cone_project(h.cones, h.w1.p, h.w2.p, EPI_NONE, E);
    h.unpack_to_host(h.w2.p, y);
    FOS_API_END(hh)
}

int32_t fos_cg_dense(fos_handle_t hh, int64_t n, const double *A, const double *b, double *x, double tol,
                     int64_t max_iters, int64_t *iters)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    FOS_REQUIRE(n > 0 && A && b && x, "bad arguments");
    FOS_REQUIRE(h.nranks == 1, "fos_cg_dense is single-GPU");
    h.loaded = false;  // replaces whatever problem was loaded on this handle
    h.A.impl = h.matvec_impl;
    h.A.init_dense(n, n, A, n, FOS_MEM_HOST, 0, n, h.grid_ctas, h.stream);
    h.L.form = 2;
    h.L.n = n;
    h.L.m = n;
    h.L.n_pad = ru(n, PAD);
    h.L.m_pad = h.L.n_pad;
    h.L.LP = 0;
    h.L.NP = h.L.n_pad;
    h.N = n;
    h.seg.nseg = 1;
    h.seg.len[0] = n;
    h.seg.src[0] = 0;
    h.seg.dst[0] = 0;
    std::vector<ConeSeg> segs;
    segs.push_back(ConeSeg{FOS_CONE_FREE, 0, 0, n});
    h.finish_load_common(segs);
    h.pack_from_host(b, h.rhs.p);
    h.pack_from_host(x, h.sol.p);
    h.cgiter = 1;
    if (tol < 0) tol = (double)n * 2.220446049250313e-16;  // conjugategradients.jl:31 default
    h.cg_solve(tol, (int)std::min<int64_t>(max_iters, 2147483647));
    h.unpack_to_host(h.sol.p, x);
    if (iters) *iters = h.cgiter;
    FOS_API_END(hh)
}

int32_t fos_prox_cone(fos_handle_t hh, int32_t cone_type, int32_t dual, const double *x, double *y, int64_t len)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    FOS_REQUIRE(len >= 0 && (len == 0 || (x && y)), "bad arguments");
    if (len == 0) return FOS_OK;
    const int64_t NP = ru(len, PAD);
    ConeSet K;
    std::vector<ConeSeg> segs;
    segs.push_back(ConeSeg{cone_type, dual ? 1 : 0, 0, len});
    K.build(NP, segs);
    DevBuf<double> din, dout;
    din.alloc((size_t)NP);
    dout.alloc((size_t)NP);
    FOS_CUDA(cudaMemcpyAsync(din.p, x, (size_t)len * 8, cudaMemcpyHostToDevice, h.stream));
    EpiArgs E{};
    h.cone_project(K, din.p, dout.p, EPI_NONE, E);
    FOS_CUDA(cudaMemcpyAsync(y, dout.p, (size_t)len * 8, cudaMemcpyDeviceToHost, h.stream));
    FOS_CUDA(cudaStreamSynchronize(h.stream));
    FOS_API_END(hh)
}

int32_t fos_get_stream(fos_handle_t hh, uint64_t *stream_out)
{
    FOS_API_BEGIN(hh)
    FOS_REQUIRE(stream_out != nullptr, "null output");
    *stream_out = (uint64_t)(uintptr_t)hh->h.stream;
    FOS_API_END(hh)
}

int32_t fos_k1_plan(int64_t m_local, int64_t n, int32_t ctas, int32_t *dims_out, int32_t *unit_begin,
                    int64_t unit_begin_cap, int32_t *slot_base, int32_t *first_cta, int64_t band_cap)
{
    if (m_local < 1 || n < 1 || ctas < 1 || !dims_out) return fail(nullptr, FOS_ERR_INVALID, "bad plan arguments");
    try {
        K1Plan P = k1_make_plan(m_local, n, ctas);
        dims_out[0] = P.G;
        dims_out[1] = P.RT;
        dims_out[2] = P.NB;
        dims_out[3] = P.nslots;
        dims_out[4] = P.kc_last;
        if (unit_begin) {
            if (unit_begin_cap < P.G + 1) return fail(nullptr, FOS_ERR_INVALID, "unit_begin too small");
            std::copy(P.cta_unit_begin.begin(), P.cta_unit_begin.end(), unit_begin);
        }
        if (slot_base && first_cta) {
            if (band_cap < P.NB + 1) return fail(nullptr, FOS_ERR_INVALID, "band arrays too small");
            std::copy(P.band_slot_base.begin(), P.band_slot_base.end(), slot_base);
            std::copy(P.band_first_cta.begin(), P.band_first_cta.end(), first_cta);
        }
        return FOS_OK;
    } catch (const std::exception &e) {
        return fail(nullptr, FOS_ERR_INVALID, e.what());
    }
}

int32_t fos_batch_plan(int64_t m, int64_t n, int64_t *out)
{
    if (!out) return fail(nullptr, FOS_ERR_INVALID, "null output");
    try {
        const BatchGeom g = batch_geometry(m, n);
        out[0] = g.lda;
        out[1] = g.ntiles;
        out[2] = g.S;
        out[3] = g.CW;
        out[4] = g.KP;
        out[5] = g.ctas_per_sm;
        out[6] = (int64_t)g.smem_bytes;
        out[7] = g.a_stride;
        return FOS_OK;
    } catch (const Error &e) {
        return fail(nullptr, e.code, e.what());
    } catch (const std::exception &e) {
        return fail(nullptr, FOS_ERR_INVALID, e.what());
    }
}

int32_t fos_hybrid_plan(int64_t m, int64_t n, const int32_t *row_nnz, int64_t *out)
{
    try {
        FOS_REQUIRE(m > 0 && n > 0 && row_nnz && out, "bad arguments");
        const HybridPlan hp = hybrid_row_plan(row_nnz, m, n);
        out[0] = hp.use ? 1 : 0;
        out[1] = hp.r0;
        out[2] = hp.md;
        out[3] = hp.sparse_rows;
        out[4] = hp.sparse_nnz;
        return FOS_OK;
    } catch (const Error &e) {
        return fail(nullptr, e.code, e.what());
    } catch (const std::exception &e) {
        return fail(nullptr, FOS_ERR_INVALID, e.what());
    }
}

int32_t fos_time_matvec(fos_handle_t hh, int32_t nvec, int32_t reps, double *ms_per_launch, double *bytes_per_launch)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    h.require_loaded();
    FOS_REQUIRE((nvec == 1 || nvec == 2) && reps >= 1, "nvec must be 1 or 2, reps >= 1");
    FOS_REQUIRE(h.L.form == 0 || nvec == 1, "two right-hand sides need the conic form");
    const double *v = h.x.p;
    const double *X[2] = {v, v + h.L.LP};
    const double *W[2] = {v + h.L.n_pad, v + h.L.LP + h.L.n_pad};
    cudaEvent_t e0, e1;
    FOS_CUDA(cudaEventCreate(&e0));
    FOS_CUDA(cudaEventCreate(&e1));
    h.A.run(nvec, X, W, nullptr, h.stream);  // warm-up
    FOS_CUDA(cudaEventRecord(e0, h.stream));
    for (int r = 0; r < reps; r++) h.A.run(nvec, X, W, nullptr, h.stream);
    FOS_CUDA(cudaEventRecord(e1, h.stream));
    FOS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FOS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_launch) *ms_per_launch = (double)ms / reps;
    if (bytes_per_launch) *bytes_per_launch = h.A.bytes_per_pass();
    FOS_API_END(hh)
}

int32_t fos_time_psd(fos_handle_t hh, int64_t d, int64_t ncones, const double *x, double *y, int32_t reps,
                     double *ms_per_call, int32_t *sweeps)
{
    FOS_API_BEGIN(hh)
    Handle &h = hh->h;
    FOS_REQUIRE(d >= 1 && d <= 1024 && ncones >= 1 && ncones <= 4096 && reps >= 1 && x, "bad arguments");
    const int64_t plen = d * (d + 1) / 2, seglen = ru(plen, PAD);
    const int64_t NP = seglen * ncones;
    ConeSet K;
    std::vector<ConeSeg> segs;
    for (int64_t k = 0; k < ncones; k++) segs.push_back(ConeSeg{FOS_CONE_SDP, 0, k * seglen, plen});
    K.build(NP, segs);
    K.psd_warm_enabled = false;  // time COLD projections: repeating the same input would otherwise start converged
    DevBuf<double> din, dout;
    din.alloc((size_t)NP);
    dout.alloc((size_t)NP);
    FOS_CUDA(cudaMemcpy2D(din.p, (size_t)seglen * 8, x, (size_t)plen * 8, (size_t)plen * 8, (size_t)ncones,
                          cudaMemcpyHostToDevice));
    FOS_SYNC_LEGACY();
    auto project = [&]() {
        if (!K.psd.empty()) psd_project(&h, K, din.p, dout.p);
        if (!K.psd_large.empty()) psd_project_large(&h, K, din.p, dout.p);
    };
    cudaEvent_t e0, e1;
    FOS_CUDA(cudaEventCreate(&e0));
    FOS_CUDA(cudaEventCreate(&e1));
    project();  // warm-up
    FOS_CUDA(cudaEventRecord(e0, h.stream));
    for (int r = 0; r < reps; r++) project();
    FOS_CUDA(cudaEventRecord(e1, h.stream));
    FOS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FOS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_per_call) *ms_per_call = (double)ms / reps;
    if (sweeps) *sweeps = K.psd_large.empty() ? 0 : psd_large_last_sweeps(&h, K);
    if (y)
        FOS_CUDA(cudaMemcpy2D(y, (size_t)plen * 8, dout.p, (size_t)seglen * 8, (size_t)plen * 8, (size_t)ncones,
                              cudaMemcpyDeviceToHost));
    FOS_API_END(hh)
}

}  // extern "C"
