// solver.cuh -- host-side state of one solver handle: device buffers, the matrix operator
// (dense TMA / plain / sparse), the affine projection S1 (CG on the KKT operator), the cone
// set S2 and the algorithm steps.  Mirrors, on the device, the objects of the reference:
//   AffinePlusLinear  utilities/affinepluslinear.jl:58-126
//   DualConeProduct   cones.jl:114-142
//   GAPData/...       solvers/*.jl
//   HSDEStatus        problemforms/HSDE/HSDEStatus.jl
#pragma once
#include <cmath>
#include <memory>

#include "common.cuh"
#include "kernels.cuh"
#include "matvec.cuh"
#include "batch.cuh"

namespace fos {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count, bool zero = true)
    {
        release();
        n = count;
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            throw Error(FOS_ERR_NOMEM, std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) +
                                           " bytes failed: " + cudaGetErrorString(e));
        }
        if (zero) {
            // cudaMemset on device memory is asynchronous and runs on the legacy default stream, which does not
            // order against the handle's non-blocking stream: wait for it, or a kernel enqueued right after the
            // allocation can be overtaken by (and wiped out by) the zero fill
            FOS_CUDA(cudaMemset(p, 0, count * sizeof(T)));
            FOS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
        }
    }
    void upload(const std::vector<T> &h)
    {
        alloc(h.size(), false);
        if (!h.empty()) {
            FOS_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
            FOS_SYNC_LEGACY();
        }
    }
};

struct Stats {
    int64_t total_cg = 0;
    int64_t total_passes = 0;
    int64_t launches = 0;
};

// ---- NCCL through dlopen: no link-time dependency; binds to whichever libnccl.so.2 the
// process already has loaded (torch's bundled one when driven from Python). ---------------
struct NcclId {
    char bytes[FOS_COMM_ID_BYTES];
};
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi &nccl_api();  // loads on first use, throws FOS_ERR_COMM if unavailable

// =======================================================================================
// MatOp: the m x n matrix A on the device and its fused dual mat-vec
// =======================================================================================
struct HybridPlan {
    bool use = false;
    int64_t r0 = 0, md = 0;  // dense block = rows [r0, r0 + md)
    int64_t sparse_rows = 0, sparse_nnz = 0;
};
HybridPlan hybrid_row_plan(const int32_t *row_nnz, int64_t m, int64_t n);  // host only

struct MatOp {
    int kind = 0;  // 0 none, 1 dense, 2 sparse, 3 hybrid (a dense row block [row_begin, row_begin + m_local) + sparse rest)
    int64_t m = 0, n = 0, n_pad = 0, m_pad = 0;
    int64_t m_local = 0, row_begin = 0, m_pad_local = 0;
    // dense
    DevBuf<double> A_own;
    const double *A = nullptr;
    int64_t lda = 0;
    CUtensorMap tmap;
    K1Plan plan;
    DevBuf<int32_t> d_unit_begin, d_slot_base, d_first_cta;
    DevBuf<double> rowpart, colpart;
    // SM-bound work ranges sized to the measured speed of every SM (calibrate(), option "k1_balance")
    bool by_smid = false;
    int k1_balance = 0;  // measured: no gain (the pass is bound by HBM, not by the slowest SM) -> opt-in
    DevBuf<unsigned int> d_claim;            // [G] claims + [G] epoch / exit ticket
    DevBuf<unsigned long long> d_cta_cycles; // [G]
    DevBuf<int32_t> d_cta_smid;              // [G]
    std::vector<double> sm_weight;           // relative speed of SM g (mean 1); empty = uniform
    double calib_spread[2] = {0.0, 0.0};     // (max - min) / mean of the per-SM times before / after balancing
    void upload_plan();
    void calibrate(cudaStream_t st);
    // plain path + sparse path: complete results
    DevBuf<double> full_ax, full_atw, scratch;
    DevBuf<int32_t> d_one_band;  // {0, 1}
    int nchunk = 0;
    // sparse
    DevBuf<int32_t> csr_ptr, csr_idx, csc_ptr, csc_idx;
    DevBuf<double> csr_val, csc_val;
    int64_t nnz = 0;
    // multi-GPU exchange
    int nranks = 1, rank = 0;
    void *comm = nullptr;
    DevBuf<double> xbuf;
    // fused exchange over NVLink peer memory (CUDA IPC, one process per GPU): k1_exchange_p2p
    DevBuf<unsigned char> p2p_region;
    DevBuf<unsigned int> p2p_local;  // [0] epoch, [1..2] tickets
    P2PView p2p{};
    bool p2p_on = false;
    std::vector<void *> p2p_opened;
    void p2p_export(uint8_t *handle_out);
    void p2p_import(const uint8_t *handles);
    void p2p_close();

    int impl = 0;  // 0 TMA, 1 plain
    int num_sms = 148;
    Stats *stats = nullptr;
    // in-situ timing of the mat-vec launches (bench.py roofline): CUDA events around every launch
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_nv;
    size_t ev_used = 0;
    double prof_ms[3] = {0, 0, 0};
    int64_t prof_n[3] = {0, 0, 0};
    int64_t prof_skipped = 0;
    void prof_begin(int NV, cudaStream_t st);
    void prof_end(cudaStream_t st);
    void prof_collect();  // call after a stream synchronisation
    void prof_reset();
    ~MatOp();

    void init_dense(int64_t m_, int64_t n_, const double *Asrc, int64_t lda_src, int location, int64_t row_begin_,
                    int64_t row_count, int grid_ctas, cudaStream_t st);
    void init_sparse(int64_t m_, int64_t n_, const int64_t *colptr, const int64_t *rowval, const double *nzval,
                     int64_t base, cudaStream_t st);
    // Hybrid row storage ("hybrid_rows"): like init_dense on the whole matrix, then the rows are classified on the
    // device; the contiguous range that holds every row with more than n/8 non-zeros stays dense (K1 streams only
    // that block), the other non-empty rows are kept as CSR + CSC.  Falls back to plain dense storage (kind 1) when
    // that would not save at least 2 % of the bytes of a pass.  Single rank only.
    // auto_only: keep the dense layout unless the bytes saved per pass are worth the three extra launches per pass
    // (>= 64 MB, i.e. >= ~10 us of HBM time: config 3 saves 3.2 GB, a 24000 x 4000 problem would lose)
    void init_hybrid(int64_t m_, int64_t n_, const double *Asrc, int64_t lda_src, int location, int grid_ctas,
                     cudaStream_t st, bool auto_only = false);
    int64_t hyb_sparse_rows = 0;  // rows kept as CSR / CSC (kind 3)
    double bytes_per_pass() const;
    // Runs one pass; X[v] have n_pad entries, W[v] have m_pad entries (global rows).
    // defer_exchange: with the peer-memory exchange enabled, return the LOCAL partial view and leave the
    // exchange to the consumer kernel (k_cg_tail_hsde<true>)
    MVView run(int NV, const double *const *X, const double *const *W, const int32_t *skip, cudaStream_t st,
               bool defer_exchange = false);

    template <int NV>
    MVView run_t(const double *const *X, const double *const *W, const int32_t *skip, cudaStream_t st,
                 bool defer_exchange);
    MVView view_full(int NV, const double *ax, const double *atw) const;
};

// =======================================================================================
// ConeSet: a product of cones over a padded vector (DualConeProduct / ConeProduct)
// =======================================================================================
struct ConeSeg {
    int32_t type;  // FOS_CONE_*
    int32_t dual;  // proxDual!
    int64_t off;   // padded offset
    int64_t len;
};
struct PsdCone {
    int64_t off;
    int32_t d;
    int32_t dual;
};
struct ConeSet {
    int64_t NP = 0;
    DevBuf<uint8_t> ops;
    DevBuf<int32_t> cone_of;
    DevBuf<SocCone> soc;
    DevBuf<double2> box;  // IndBox bounds, indexed by cone_of[e] where ops[e] == OP_BOX
    std::vector<double2> h_box;
    std::vector<uint8_t> h_ops_keep;
    std::vector<int32_t> h_cone_of_keep;
    void set_box(int64_t off, int64_t len, double lo, double hi);
    DevBuf<SocScale> soc_scale;
    DevBuf<int32_t> chunk_cone;
    DevBuf<double> chunk_sum;
    DevBuf<unsigned int> counter;
    int nsoc = 0, nchunks = 0;
    std::vector<PsdCone> psd;        // d <= PSD_SMEM_MAX_D: one CTA per cone, two-sided Jacobi in shared memory
    std::vector<PsdCone> psd_large;  // larger: cooperative one-sided block Jacobi (psd_large.cu)
    DevBuf<PsdCone> d_psd, d_psd_large;
    DevBuf<double> psd_work;
    DevBuf<uint8_t> psd_ctl;
    DevBuf<double> psd_vstore;   // eigenvector bases of the previous projection (warm start of K5L)
    bool psd_warm = false, psd_warm_enabled = true;
    // psd is sorted by order: [0, psd_n16) one warp per cone, [psd_n16, psd_n48) 128 threads, the rest 512 (psd.cu)
    int psd_max_d = 0, psd_n16 = 0, psd_n48 = 0, psd_dmax16 = 0, psd_dmax48 = 0;
    bool fusable = true;  // every cone projects entry by entry from (x_e, per-cone scalars): RelaxArgs may feed it
    void build(int64_t NP_, const std::vector<ConeSeg> &segs);
};

struct Handle;

// batch mode (batch.cu): B problems of identical shape, one persistent CTA per problem
struct BatchSolver {
    Handle *h = nullptr;
    bool loaded = false;
    int64_t B = 0, N = 0;
    Lay L{};
    SegMap seg{};
    int64_t lda = 0, a_stride = 0;
    int ntiles = 0, S = 0, CW = 0, KP = 1, ctas_per_sm = 1;
    size_t smem_bytes = 0;
    int grid_ctas = 0;
    DevBuf<double> dA, db, dc, dnb, dncn, vec, drecs, dtol, d_sval, d_cval;
    DevBuf<int32_t> d_drow, d_cnt, d_sptr, d_sid, d_scol, d_cptr, d_cid, d_crow;
    int32_t NS = 1, NCS = 1, NZ = 1;
    int64_t dense_rows_total = 0, sparse_nnz_total = 0;
    bool hybrid = true;  // rows with <= n/8 non-zeros as CSR/CSC instead of dense tiles
    DevBuf<BatchCtl> dctl;
    DevBuf<unsigned int> counter;
    ConeSet cones;
    int rec_cap = 0, tol_n = 0;
    double tol_floor = 0.0;
    double last_eps = 1e-5;  // eps of the last run (the forced final check of finish uses it)

    void load(Handle *h_, int64_t B_, int64_t m, int64_t n, const double *A, int64_t lda_src, int64_t pstride_src,
              int location, const double *b, const double *c, int64_t nc1, const int32_t *t1, const int64_t *l1,
              int64_t nc2, const int32_t *t2, const int64_t *l2);
    void ensure_recs(int cap);
    void set_algorithm();
    void begin_solve();
    void set_initial_iterate();
    void set_vector(int which, const double *z, int64_t pb0, int64_t count);
    void get_vector(int which, double *z, int64_t pb0, int64_t count);
    void launch(int64_t i_start, int64_t n_iters, int64_t checki, double eps, bool do_run, bool do_finish);
    void collect(int64_t *iters_done, int32_t *status, double *records, int64_t rec_cap_host, int64_t *n_rec,
                 int64_t *cgiter_total);
    double bytes_per_pass() const;
};

struct Handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    std::string err;
    // options
    int matvec_impl = 0, grid_ctas = 0, cg_batch = 0, fuse_rhs = 1, batch_ctas = 0, fuse_tail = 1, batch_hybrid = 1, tail_blocks = 0, tail_flag_mode = 0, hybrid_rows = 2;
    // problem
    bool loaded = false;
    Lay L{};
    SegMap seg{};
    int64_t N = 0;
    MatOp A;
    DevBuf<double> d_b, d_c, d_q, d_bhat;  // HSDE: b (m_pad), c (n_pad).  plain: q (n_pad), b (m_pad)
    double nb = 0, ncn = 0;
    double beta = 1.0;
    bool decreasing = true;
    // S1
    DevBuf<double> rhs, sol, r, p, Ap;
    int64_t s1_calls = 1, cgiter = 0;
    bool firstrun = true;
    bool warn_maxit = false;
    // algorithm
    int alg = FOS_ALG_GAP;
    double alpha = 0.8, alpha1 = 1.8, alpha2 = 1.8, betaA = 0.0;
    int64_t iproj = 100;
    int64_t lsinterval = 0;  // LineSearchWrapper (wrappers/linesearch.jl); 0 = no wrapper
    double fista_t = 1.0;
    DevBuf<double> x, tmp1, tmp2, proj, fy, fxold, dp, dq, dy, w1, w2, w3, prev;
    ConeSet cones;
    // control
    DevBuf<Ctrl> d_ctrl;
    Ctrl *h_ctrl = nullptr;  // pinned
    DevBuf<double> red_partials;
    DevBuf<unsigned int> red_counter;
    RedBuf rb{};
    DevBuf<unsigned int> gb_ctr;  // fused CG tail: grid-barrier counters {count, base, exit ticket}
    DevBuf<double> gb_part;
    GridBar gbar{};
    DevBuf<unsigned long long> tail_trace;  // option "tail_trace": [3][16] phase cycle sums of k_cg_tail_hsde
    int tail_occ[2] = {0, 0};  // co-resident blocks per SM of k_cg_tail_hsde<false / true> (queried once)
    int tail_grid = 0;
    DevBuf<double> d_recs;
    int rec_cap = 0;
    DevBuf<double> d_stage;
    double *h_stage = nullptr;  // pinned
    size_t h_stage_n = 0;
    int status = FOS_STATUS_CONTINUE;
    bool checked = false;
    Stats stats;
    // comm
    int rank = 0, nranks = 1;
    void *comm = nullptr;
    // run parameters remembered for finish()
    int64_t cur_i = 0, cur_checki = 100;
    double cur_eps = 1e-5;
    // graph path (option "use_graphs", default 1): one CUDA graph per outer iteration of GAP / DR / AP / GAPA on the
    // conic form -- [k_iter_begin, K1, initial residual] -> WHILE(!done){K1, fused CG tail} -> [cone projection fused
    // with both relaxations and the averaging] (+ [K1, k6] on check iterations) -- launched back to back with no host
    // synchronisation between iterations; the CG loop is a conditional graph node driven by the tail's stop test
    struct IterGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        int fixed_launches = 0;  // kernels outside the CG loop
    };
    IterGraph graphs[2];  // [0] plain iteration, [1] iteration with a status check
    int use_graphs = 1;
    bool capturing = false;
    double graph_eps = -1.0;
    cudaStream_t stream2 = nullptr;  // captures the body of the WHILE node
    DevBuf<double> tol_table;        // max(0.2^sqrt(i), l*eps) for S1.i = 0 .. TOL_TABLE_N-1 (host pow, like the legacy path)
    static constexpr int TOL_TABLE_N = 512;
    bool graph_ok() const;
    void drop_graphs();
    void build_iter_graph(bool with_check);
    int64_t run_graph(int64_t i_start, int64_t n_iters);
    // batch mode
    std::unique_ptr<BatchSolver> batch;
    // direct = true (direct.cu): W = (I + Q Q')^-1 on the device, streamed by its own MatOp
    bool direct = false;
    DevBuf<double> Winv;
    std::unique_ptr<MatOp> Wop;
    void set_direct(bool on);
    void s1_prox_direct(const double *xin);

    ~Handle();
    void create(int dev);
    void require_loaded() const { FOS_REQUIRE(loaded, "no problem loaded on this handle"); }
    // loading
    void finish_load_common(const std::vector<ConeSeg> &segs);
    void load_conic(int64_t m, int64_t n, const double *b, const double *c, int64_t nc1, const int32_t *t1,
                    const int64_t *l1, int64_t nc2, const int32_t *t2, const int64_t *l2);
    void load_affine(int64_t am, int64_t an, const double *b, const double *q, int32_t beta_, int32_t decr,
                     int64_t nc, const int32_t *t, const int64_t *l);
    // vectors
    int vgrid(int64_t len) const;
    void ensure_stage(size_t n);
    void pack_from_host(const double *z, double *dst);
    void unpack_to_host(const double *src, double *z);
    // operators
    MVView kkt_pass(const double *v, const int32_t *skip, bool defer_exchange = false);
    void kkt_mul(const double *in, double *out);
    void q_mul(const double *Bp, double *Yp, bool transpose);
    void s1_prox(const double *xin);
    void cg_solve(double tol, int max_iters, const double *x0 = nullptr, const double *rhs_ = nullptr);
    void cg_enqueue_iteration(unsigned long long cond = 0ull);
    void sol_scaled_to(double *dst);
    void cone_project(ConeSet &K, const double *in, double *projbuf, int epi, const EpiArgs &E);
    void check(const double *z, int64_t i, bool override_);
    void sync_ctrl();
    // algorithm steps
    void set_algorithm(int alg_, double a, double a1, double a2, double bA, int64_t ip);
    void step(int64_t i);
    void getsol();
    void begin_solve();
    int64_t run(int64_t i_start, int64_t n_iters, int64_t checki, double eps, double *records, int64_t rec_cap_host,
                int64_t *n_rec, double *trace);
    void finish(double *guess, double *record, int64_t *n_rec);
    void ensure_recs(int cap);
};

#define FOS_LAUNCH(h, kernel, grid, block, smem, ...)                  \
    do {                                                               \
        kernel<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__); \
        (h)->stats.launches++;                                         \
    } while (0)

constexpr int PSD_SMEM_MAX_D = 112;  // largest cone whose S and V fit one SM's shared memory
constexpr int PSD_WARP_MAX_D = 16;   // up to this order a cone is projected by ONE WARP (eight cones per CTA) ...
extern int g_psd_warp_max_d;         // ... unless the option "psd_warp_max_d" lowers it (process-wide; 0 = never)
void psd_project(Handle *h, ConeSet &K, const double *in, double *projbuf);        // K5, psd.cu
void psd_project_large(Handle *h, ConeSet &K, const double *in, double *projbuf);  // K5, psd_large.cu
int psd_large_last_sweeps(Handle *h, ConeSet &K);

}  // namespace fos
