// solver.cu -- the hot loop on the device: affine projection by CG on the KKT operator,
// cone projection, relaxation, residual checks.  See solver.cuh for the object map.
#include <algorithm>

#include "solver.cuh"

namespace fos {

int g_psd_warp_max_d = 16;  // at most 16: the 128-thread kernel takes over from there (psd.cu)

// =======================================================================================
// handle lifecycle
// =======================================================================================
void Handle::create(int dev)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(FOS_ERR_CUDA, std::string("no CUDA device available (this library has no CPU path): ") +
                                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    FOS_REQUIRE(dev >= 0 && dev < count, "device index out of range");
    device = dev;
    FOS_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    FOS_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        throw Error(FOS_ERR_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) +
                                      std::to_string(prop.minor) + "; this library is built for sm_100a only");
    num_sms = prop.multiProcessorCount;
    FOS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FOS_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
    d_ctrl.alloc(1);
    FOS_CUDA(cudaMallocHost((void **)&h_ctrl, sizeof(Ctrl)));
    memset(h_ctrl, 0, sizeof(Ctrl));
    red_partials.alloc((size_t)RED_SLOTS * MAX_PARTIALS);
    red_counter.alloc(1);
    rb.partials = red_partials.p;
    rb.counter = red_counter.p;
    A.num_sms = num_sms;
    A.stats = &stats;
    gb_ctr.alloc(96);
    gb_part.alloc((size_t)2 * 2 * num_sms * 8);
    gbar.count = gb_ctr.p;
    gbar.base = gb_ctr.p + 1;
    gbar.exit_ticket = gb_ctr.p + 2;
    gbar.flag = gb_ctr.p + 64;
    gbar.part = gb_part.p;
    gbar.trace = nullptr;
}

Handle::~Handle()
{
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    if (comm) {
        try {
            nccl_api().CommDestroy(comm);
        } catch (...) {
        }
    }
    drop_graphs();
    if (h_ctrl) cudaFreeHost(h_ctrl);
    if (h_stage) cudaFreeHost(h_stage);
    if (stream2) cudaStreamDestroy(stream2);
    if (stream) cudaStreamDestroy(stream);
}

int Handle::vgrid(int64_t len) const
{
    int64_t g = (len + VBLOCK - 1) / VBLOCK;
    int64_t cap = std::min<int64_t>(4 * (int64_t)num_sms, MAX_PARTIALS);
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

void Handle::ensure_stage(size_t n)
{
    if (h_stage_n >= n) return;
    if (h_stage) cudaFreeHost(h_stage);
    h_stage = nullptr;
    FOS_CUDA(cudaMallocHost((void **)&h_stage, n * sizeof(double)));
    h_stage_n = n;
    d_stage.alloc(n);
}

void Handle::pack_from_host(const double *z, double *dst)
{
    ensure_stage((size_t)N);
    memcpy(h_stage, z, (size_t)N * sizeof(double));
    FOS_CUDA(cudaMemcpyAsync(d_stage.p, h_stage, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, stream));
    FOS_LAUNCH(this, k_pack, vgrid(L.NP), VBLOCK, 0, seg, d_stage.p, dst, L.NP);
    FOS_CUDA(cudaStreamSynchronize(stream));
}

void Handle::unpack_to_host(const double *src, double *z)
{
    ensure_stage((size_t)N);
    FOS_LAUNCH(this, k_unpack, vgrid(N), VBLOCK, 0, seg, src, d_stage.p, N);
    FOS_CUDA(cudaMemcpyAsync(h_stage, d_stage.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, stream));
    FOS_CUDA(cudaStreamSynchronize(stream));
    memcpy(z, h_stage, (size_t)N * sizeof(double));
}

void Handle::sync_ctrl()
{
    FOS_CUDA(cudaMemcpyAsync(h_ctrl, d_ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
    if (A.p2p_on) FOS_CUDA(cudaMemcpyAsync(&h_ctrl->p2p_error, A.p2p.error, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    FOS_CUDA(cudaStreamSynchronize(stream));
    if (A.p2p_on && h_ctrl->p2p_error != 0)
        throw Error(FOS_ERR_COMM, "peer-memory exchange timed out: a rank did not reach the matching pass over A "
                                  "(every rank must make the same calls in the same order)");
    A.prof_collect();
}

void Handle::ensure_recs(int cap)
{
    if (cap <= rec_cap) return;
    drop_graphs();  // the record buffer is a kernel argument of the captured status check
    d_recs.alloc((size_t)cap * FOS_REC_LEN);
    rec_cap = cap;
}

// =======================================================================================
// cone set
// =======================================================================================
void ConeSet::build(int64_t NP_, const std::vector<ConeSeg> &segs)
{
    NP = NP_;
    std::vector<uint8_t> h_ops((size_t)NP, OP_ZERO);
    std::vector<int32_t> h_cone_of((size_t)NP, 0);
    std::vector<SocCone> h_soc;
    std::vector<int32_t> h_chunk_cone;
    psd.clear();
    psd_large.clear();
    psd_max_d = 0;
    fusable = true;
    for (const ConeSeg &s : segs) {
        // cones whose projection reads NEIGHBOURING input entries (or runs its own kernels) cannot take their
        // input from the fused relaxation (RelaxArgs): rotated SOC, exponential cones, PSD
        if (s.len > 0 && (s.type == FOS_CONE_SOCROT || s.type == FOS_CONE_EXPPRIMAL || s.type == FOS_CONE_EXPDUAL ||
                          s.type == FOS_CONE_SDP))
            fusable = false;
        FOS_REQUIRE(s.off >= 0 && s.off + s.len <= NP && s.len >= 0, "cone segment out of range");
        uint8_t op = OP_ZERO;
        switch (s.type) {
        case FOS_CONE_FREE: op = s.dual ? OP_ZERO : OP_COPY; break;  // cones.jl:100
        case FOS_CONE_ZERO: op = s.dual ? OP_COPY : OP_ZERO; break;  // cones.jl:98
        case FOS_CONE_NONNEG: op = OP_MAX0; break;                   // cones.jl:101
        case FOS_CONE_NONPOS: op = OP_MIN0; break;                   // cones.jl:102
        case FOS_CONE_SOC:
        case FOS_CONE_SOCROT: {
            if (s.len == 0) continue;
            const bool rot = s.type == FOS_CONE_SOCROT;
            FOS_REQUIRE(!rot || s.len >= 2, "a rotated SOC needs at least 2 entries");
            SocCone c;
            c.head = s.off;
            c.len = s.len;
            c.dual = s.dual;
            c.rot = rot ? 1 : 0;
            c.chunk0 = (int32_t)h_chunk_cone.size();
            int64_t tail = s.len - (rot ? 2 : 1);
            c.nchunk = (int32_t)std::max<int64_t>(1, (tail + SOC_CHUNK - 1) / SOC_CHUNK);
            for (int k = 0; k < c.nchunk; k++) h_chunk_cone.push_back((int32_t)h_soc.size());
            const int64_t nhead = rot ? 2 : 1;
            for (int64_t k = 0; k < s.len; k++) {
                uint8_t o;
                if (rot) o = k == 0 ? (s.dual ? OP_RSOCD_H0 : OP_RSOC_H0) : k == 1 ? (s.dual ? OP_RSOCD_H1 : OP_RSOC_H1)
                                                                                  : (s.dual ? OP_RSOCD_TAIL : OP_RSOC_TAIL);
                else o = k < nhead ? (s.dual ? OP_SOCD_HEAD : OP_SOC_HEAD) : (s.dual ? OP_SOCD_TAIL : OP_SOC_TAIL);
                h_ops[(size_t)(s.off + k)] = o;
                h_cone_of[(size_t)(s.off + k)] = (int32_t)h_soc.size();
            }
            h_soc.push_back(c);
            continue;
        }
        case FOS_CONE_EXPPRIMAL:
        case FOS_CONE_EXPDUAL: {
            // cones.jl:12-13; a cone entry may hold several 3-blocks (MathProgBase passes one triple per cone)
            FOS_REQUIRE(s.len % 3 == 0, "exponential cones have 3 entries");
            const uint8_t o = s.type == FOS_CONE_EXPPRIMAL ? (s.dual ? OP_EXPP_D : OP_EXPP) : (s.dual ? OP_EXPD_D : OP_EXPD);
            for (int64_t k = 0; k < s.len; k++) {
                h_ops[(size_t)(s.off + k)] = o;
                h_cone_of[(size_t)(s.off + k)] = (int32_t)(k % 3);
            }
            continue;
        }
        case FOS_CONE_SDP: {
            if (s.len == 0) continue;
            int64_t d = (int64_t)std::llround(std::sqrt(0.25 + 2.0 * (double)s.len) - 0.5);
            FOS_REQUIRE(d * (d + 1) / 2 == s.len, "SDP cone length is not d(d+1)/2");
            PsdCone pc;
            pc.off = s.off;
            pc.d = (int32_t)d;
            pc.dual = s.dual;
            if (d <= PSD_SMEM_MAX_D) {
                psd.push_back(pc);
                psd_max_d = std::max<int>(psd_max_d, (int)d);
            } else {
                FOS_REQUIRE(d <= 1024, "SDP cones larger than 1024 x 1024 are not supported");
                psd_large.push_back(pc);
            }
            for (int64_t k = 0; k < s.len; k++) h_ops[(size_t)(s.off + k)] = OP_PRE;
            continue;
        }
        default:
            throw Error(FOS_ERR_UNSUPPORTED, "unknown cone type " + std::to_string(s.type));
        }
        for (int64_t k = 0; k < s.len; k++) h_ops[(size_t)(s.off + k)] = op;
    }
    ops.upload(h_ops);
    cone_of.upload(h_cone_of);
    h_ops_keep = h_ops;
    h_cone_of_keep = h_cone_of;
    h_box.clear();
    box.release();
    nsoc = (int)h_soc.size();
    nchunks = (int)h_chunk_cone.size();
    if (nsoc > 0) {
        soc.upload(h_soc);
        chunk_cone.upload(h_chunk_cone);
        soc_scale.alloc((size_t)nsoc);
        chunk_sum.alloc((size_t)nchunks);
    }
    counter.alloc(1);
    if (!psd.empty()) {
        // sorted by order: the projection kernel gives a cone 32, 128 or 512 threads depending on d (psd.cu)
        std::stable_sort(psd.begin(), psd.end(), [](const PsdCone &a, const PsdCone &b) { return a.d < b.d; });
        psd_n16 = psd_n48 = 0;
        psd_dmax16 = psd_dmax48 = psd_max_d = 0;
        for (const PsdCone &c : psd) {
            if (c.d <= g_psd_warp_max_d) {
                psd_n16++;
                psd_dmax16 = std::max<int>(psd_dmax16, c.d);
            } else if (c.d <= 48) {
                psd_dmax48 = std::max<int>(psd_dmax48, c.d);
            } else {
                psd_max_d = std::max<int>(psd_max_d, c.d);
            }
            if (c.d <= 48) psd_n48++;
        }
        if (psd_n48 < psd_n16) psd_n48 = psd_n16;
        d_psd.upload(psd);
    }
    if (!psd_large.empty()) d_psd_large.upload(psd_large);
}

// IndBox(lo, hi) over padded entries [off, off+len): overrides whatever elementwise cone was there
void ConeSet::set_box(int64_t off, int64_t len, double lo, double hi)
{
    FOS_REQUIRE(off >= 0 && len >= 0 && off + len <= NP, "box range out of bounds");
    FOS_REQUIRE(!(lo > hi), "IndBox needs lo <= hi");
    for (int64_t k = 0; k < len; k++) {
        const uint8_t o = h_ops_keep[(size_t)(off + k)];
        FOS_REQUIRE(o <= OP_MIN0 || o == OP_BOX, "a box cannot overlap a SOC / SDP / exponential cone");
        h_ops_keep[(size_t)(off + k)] = OP_BOX;
        h_cone_of_keep[(size_t)(off + k)] = (int32_t)h_box.size();
    }
    h_box.push_back(make_double2(lo, hi));
    ops.upload(h_ops_keep);
    cone_of.upload(h_cone_of_keep);
    box.upload(h_box);
}

void Handle::cone_project(ConeSet &K, const double *in, double *projbuf, int epi, const EpiArgs &E)
{
    const RelaxArgs R{};
    if (K.nsoc > 0)
        FOS_LAUNCH(this, k4_soc_norms<false>, K.nchunks, VBLOCK, 0, in, K.soc.p, K.nsoc, K.chunk_cone.p,
                   K.chunk_sum.p, K.soc_scale.p, K.counter.p, R, d_ctrl.p);
    if (!K.psd.empty()) psd_project(this, K, in, projbuf);
    if (!K.psd_large.empty()) psd_project_large(this, K, in, projbuf);
    const int g = vgrid(K.NP);
#define CONE_CASE(EPI)                                                                                            \
    case EPI:                                                                                                     \
        FOS_LAUNCH(this, k4_cone_apply<EPI>, g, VBLOCK, 0, K.NP, in, projbuf, K.ops.p, K.cone_of.p, K.soc_scale.p, \
                   K.soc.p, K.box.p, E, d_ctrl.p, rb, R);                                                         \
        break;
    switch (epi) {
        CONE_CASE(EPI_NONE)
        CONE_CASE(EPI_GAP)
        CONE_CASE(EPI_GAPA)
        CONE_CASE(EPI_FISTA)
        CONE_CASE(EPI_DYKSTRA)
        CONE_CASE(EPI_GAPP_PROJ)
        CONE_CASE(EPI_LS)
        CONE_CASE(EPI_LSW)
    default: throw Error(FOS_ERR_INVALID, "bad epilogue");
    }
#undef CONE_CASE
}

// =======================================================================================
// loading
// =======================================================================================
static void add_cones(std::vector<ConeSeg> &segs, int64_t base, int64_t nc, const int32_t *t, const int64_t *l,
                      int dual)
{
    int64_t off = base;
    for (int64_t k = 0; k < nc; k++) {
        ConeSeg s;
        s.type = t[k];
        s.dual = dual;
        s.off = off;
        s.len = l[k];
        FOS_REQUIRE(l[k] >= 0, "negative cone length");
        segs.push_back(s);
        off += l[k];
    }
}

void Handle::finish_load_common(const std::vector<ConeSeg> &segs)
{
    const size_t NP = (size_t)L.NP;
    for (DevBuf<double> *b : {&rhs, &sol, &r, &p, &Ap, &x, &tmp1, &tmp2, &proj, &fy, &fxold, &dp, &dq, &dy, &w1, &w2,
                              &w3, &prev})
        b->alloc(NP);
    cones.build(L.NP, segs);
    drop_graphs();
    {
        // the tolerance schedule of affinepluslinear.jl:108-112 for the graph path, from the same host pow() the
        // legacy path uses; beyond the table the schedule sits on its floor an*eps
        const double an_ = (double)(L.form == 0 ? (L.n + L.m + 1) : L.n);
        const double floor_ = an_ * 2.220446049250313e-16;
        std::vector<double> tt((size_t)TOL_TABLE_N);
        for (int k = 0; k < TOL_TABLE_N; k++)
            tt[(size_t)k] = decreasing ? std::max(std::pow(0.2, std::sqrt((double)k)), floor_) : floor_;
        tt[(size_t)TOL_TABLE_N - 1] = floor_;
        tol_table.upload(tt);
    }
    s1_calls = 1;
    cgiter = 0;
    firstrun = true;
    warn_maxit = false;
    direct = false;
    Wop.reset();
    Winv.release();
    lsinterval = 0;
    status = FOS_STATUS_CONTINUE;
    checked = false;
    stats = Stats();
    FOS_CUDA(cudaMemset(d_ctrl.p, 0, sizeof(Ctrl)));
    FOS_CUDA(cudaStreamSynchronize(cudaStreamLegacy));  // see DevBuf::alloc
    ensure_recs(256);
    ensure_stage((size_t)N);
    loaded = true;
    set_algorithm(FOS_ALG_GAP, 0.8, 1.8, 1.8, 0.0, 100);
}

// HSDE(model) (problemforms/HSDE/HSDE.jl:7-29, indirect branch) on a matrix already in `A`
void Handle::load_conic(int64_t m, int64_t n, const double *b, const double *c, int64_t nc1, const int32_t *t1,
                        const int64_t *l1, int64_t nc2, const int32_t *t2, const int64_t *l2)
{
    int64_t s1 = 0, s2 = 0;
    for (int64_t k = 0; k < nc1; k++) s1 += l1[k];
    for (int64_t k = 0; k < nc2; k++) s2 += l2[k];
    FOS_REQUIRE(s1 == m, "constraint cones do not cover 1:m (cones.jl:66-72)");
    FOS_REQUIRE(s2 == n, "variable cones do not cover 1:n (cones.jl:66-72)");
    L.form = 0;
    L.n = n;
    L.m = m;
    L.n_pad = ru(n, PAD);
    L.m_pad = ru(m, PAD);
    L.LP = L.n_pad + L.m_pad + PAD;
    L.NP = 2 * L.LP;
    const int64_t l = m + n + 1;
    N = 2 * l;
    seg.nseg = 6;
    const int64_t lens[6] = {n, m, 1, n, m, 1};
    const int64_t srcs[6] = {0, n, n + m, l, l + n, l + n + m};
    const int64_t dsts[6] = {0, L.n_pad, L.n_pad + L.m_pad, L.LP, L.LP + L.n_pad, L.LP + L.n_pad + L.m_pad};
    for (int k = 0; k < 6; k++) {
        seg.len[k] = lens[k];
        seg.src[k] = srcs[k];
        seg.dst[k] = dsts[k];
    }
    std::vector<double> hb((size_t)L.m_pad, 0.0), hc((size_t)L.n_pad, 0.0);
    double sb = 0, sc = 0;
    for (int64_t i = 0; i < m; i++) {
        hb[(size_t)i] = b[i];
        sb += b[i] * b[i];
    }
    for (int64_t j = 0; j < n; j++) {
        hc[(size_t)j] = c[j];
        sc += c[j] * c[j];
    }
    nb = std::sqrt(sb);
    ncn = std::sqrt(sc);
    d_b.upload(hb);
    d_c.upload(hc);
    beta = 1.0;
    decreasing = true;  // HSDE.jl:22
    // DualConeProduct (cones.jl:114-142): K2 x K1* x R+ x K2* x K1 x R+
    std::vector<ConeSeg> segs;
    add_cones(segs, 0, nc2, t2, l2, 0);
    add_cones(segs, L.n_pad, nc1, t1, l1, 1);
    segs.push_back(ConeSeg{FOS_CONE_NONNEG, 0, L.n_pad + L.m_pad, 1});
    add_cones(segs, L.LP, nc2, t2, l2, 1);
    add_cones(segs, L.LP + L.n_pad, nc1, t1, l1, 0);
    segs.push_back(ConeSeg{FOS_CONE_NONNEG, 0, L.LP + L.n_pad + L.m_pad, 1});
    finish_load_common(segs);
}

// Feasibility(AffinePlusLinear(A,b,q,beta), ConeProduct, an+am)
void Handle::load_affine(int64_t am, int64_t an, const double *b, const double *q, int32_t beta_, int32_t decr,
                         int64_t nc, const int32_t *t, const int64_t *l)
{
    FOS_REQUIRE(beta_ == 1 || beta_ == -1, "beta must be 1 or -1 (affinepluslinear.jl:73)");
    int64_t s1 = 0;
    for (int64_t k = 0; k < nc; k++) s1 += l[k];
    FOS_REQUIRE(s1 == am + an, "cones do not cover 1:(an+am)");
    L.form = 1;
    L.n = an;
    L.m = am;
    L.n_pad = ru(an, PAD);
    L.m_pad = ru(am, PAD);
    L.LP = 0;
    L.NP = L.n_pad + L.m_pad;
    N = an + am;
    seg.nseg = 2;
    seg.len[0] = an; seg.src[0] = 0;  seg.dst[0] = 0;
    seg.len[1] = am; seg.src[1] = an; seg.dst[1] = L.n_pad;
    std::vector<double> hq((size_t)L.n_pad, 0.0);
    for (int64_t j = 0; j < an; j++) hq[(size_t)j] = q ? q[j] : 0.0;
    d_q.upload(hq);
    beta = (double)beta_;
    decreasing = decr != 0;
    // cones over [x; z]: split the contiguous cone list at the x/z boundary of the padded layout.
    std::vector<ConeSeg> segs;
    int64_t off = 0;
    for (int64_t k = 0; k < nc; k++) {
        FOS_REQUIRE(l[k] >= 0, "negative cone length");
        const int64_t a0 = off, a1 = off + l[k];
        if (a1 <= an) segs.push_back(ConeSeg{t[k], 0, a0, l[k]});
        else if (a0 >= an) segs.push_back(ConeSeg{t[k], 0, L.n_pad + (a0 - an), l[k]});
        else {
            FOS_REQUIRE(t[k] >= FOS_CONE_FREE && t[k] <= FOS_CONE_NONPOS,
                        "only elementwise cones may straddle the x/z boundary of the affine form");
            segs.push_back(ConeSeg{t[k], 0, a0, an - a0});
            segs.push_back(ConeSeg{t[k], 0, L.n_pad, a1 - an});
        }
        off = a1;
    }
    finish_load_common(segs);
    // rhs2 = b (affinepluslinear.jl:76-77)
    std::vector<double> hb((size_t)L.m_pad, 0.0);
    for (int64_t i = 0; i < am; i++) hb[(size_t)i] = b ? b[i] : 0.0;
    FOS_CUDA(cudaMemcpy(rhs.p + L.n_pad, hb.data(), (size_t)L.m_pad * 8, cudaMemcpyHostToDevice));
    FOS_SYNC_LEGACY();
    d_bhat.upload(hb);
}

// =======================================================================================
// operators
// =======================================================================================
// one pass over A feeding the KKT product of v
MVView Handle::kkt_pass(const double *v, const int32_t *skip, bool defer_exchange)
{
    if (L.form == 0) {
        const double *X[2] = {v, v + L.LP};
        const double *W[2] = {v + L.n_pad, v + L.LP + L.n_pad};
        return A.run(2, X, W, skip, stream, defer_exchange);
    }
    const double *X[1] = {v};
    const double *W[1] = {L.form == 2 ? v : v + L.n_pad};
    return A.run(1, X, W, skip, stream);
}

void Handle::kkt_mul(const double *in, double *out)
{
    MVView V = kkt_pass(in, nullptr);
    if (L.form == 0)
        FOS_LAUNCH(this, k2_kkt_hsde<K2_OUT>, vgrid(L.LP), VBLOCK, 0, L, V, in, d_c.p, d_b.p, out, nullptr, nullptr,
                   nullptr, d_ctrl.p, rb, 0);
    else
        FOS_LAUNCH(this, k2_kkt_plain<K2_OUT>, vgrid(L.NP), VBLOCK, 0, L, V, in, out, nullptr, nullptr, nullptr,
                   d_ctrl.p, rb, 0);
}

void Handle::q_mul(const double *Bp, double *Yp, bool transpose)
{
    FOS_REQUIRE(L.form == 0, "q_mul needs the HSDE conic form");
    const double *X[1] = {Bp};
    const double *W[1] = {Bp + L.n_pad};
    MVView V = A.run(1, X, W, nullptr, stream);
    FOS_LAUNCH(this, k2_q_hsde, vgrid(L.LP), VBLOCK, 0, L, V, Bp, d_c.p, d_b.p, Yp, transpose ? 1 : 0, rb);
}

void Handle::cg_enqueue_iteration(unsigned long long cond)
{
    // cond != 0: this launch pair is the body of a CUDA-graph WHILE node (run_graph): no predication needed
    const int32_t *skip = cond != 0ull ? nullptr : &d_ctrl.p->done;
    if (L.form == 0 && fuse_tail) {
        // K1 + ONE cooperative kernel for the rest of the iteration (k_cg_tail_hsde)
        const bool p2p = A.nranks > 1 && A.p2p_on;
        MVView V = kkt_pass(p.p, skip, /*defer_exchange=*/p2p);
        // identical on every rank (block i of all ranks owns the same entries); co-resident by construction
        // single GPU: one block per SM (more blocks do not shorten the latency chain).  Peer exchange: up to two
        // per SM, so that a thread gathers ONE entry (one NVLink round trip) rather than two in sequence
        const int64_t auto_blocks = p2p ? 2 * (int64_t)num_sms : (int64_t)num_sms;
        const void *fn = p2p ? (const void *)k_cg_tail_hsde<true> : (const void *)k_cg_tail_hsde<false>;
        // a cooperative grid must be co-resident: never ask for more blocks than the kernel's occupancy allows
        // (same binary and same GPU model on every rank, so the clamp is identical everywhere)
        int &occ = tail_occ[p2p ? 1 : 0];
        if (occ == 0) {
            FOS_REQUIRE(!capturing, "internal: occupancy query during graph capture");
            FOS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, VBLOCK, 0));
            FOS_REQUIRE(occ >= 1, "the fused CG tail does not fit on an SM");
        }
        const int64_t resident = (int64_t)occ * num_sms;
        const int grid = (int)std::min<int64_t>(
            std::min<int64_t>((L.LP + VBLOCK - 1) / VBLOCK, resident),
            std::min<int64_t>(tail_blocks > 0 ? tail_blocks : auto_blocks, P2P_MAX_BLOCKS));
        const double *cptr = d_c.p, *bptr = d_b.p;
        double *solp = sol.p, *rp = r.p, *pp = p.p, *App = Ap.p;
        Ctrl *cp = d_ctrl.p;
        void *args[] = {(void *)&L, (void *)&V, (void *)&A.p2p, (void *)&cptr, (void *)&bptr, (void *)&solp,
                        (void *)&rp,  (void *)&pp, (void *)&App, (void *)&cp, (void *)&gbar, (void *)&cond};
        A.prof_begin(0, stream);
        FOS_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(VBLOCK), args, 0, stream));
        A.prof_end(stream);
        stats.launches++;
        return;
    }
    MVView V = kkt_pass(p.p, skip);
    if (L.form == 0)
        FOS_LAUNCH(this, k2_kkt_hsde<K2_AP>, vgrid(L.LP), VBLOCK, 0, L, V, p.p, d_c.p, d_b.p, Ap.p, nullptr, nullptr,
                   nullptr, d_ctrl.p, rb, 1);
    else if (L.form == 1)
        FOS_LAUNCH(this, k2_kkt_plain<K2_AP>, vgrid(L.NP), VBLOCK, 0, L, V, p.p, Ap.p, nullptr, nullptr, nullptr,
                   d_ctrl.p, rb, 1);
    else
        FOS_LAUNCH(this, k2_spd<K2_AP>, vgrid(L.NP), VBLOCK, 0, L, V, p.p, Ap.p, nullptr, nullptr, nullptr, d_ctrl.p,
                   rb, 1);
    FOS_LAUNCH(this, k3_cg_update, vgrid(L.NP), VBLOCK, 0, L.NP, sol.p, r.p, p.p, Ap.p, d_ctrl.p, rb);
    FOS_LAUNCH(this, k3_cg_dir, vgrid(L.NP), VBLOCK, 0, L.NP, p.p, r.p, d_ctrl.p);
}

// conjugategradient!(sol, KKT, rhs, r, p, Ap; tol, max_iters) (conjugategradients.jl:31-55).
// Iterations are enqueued in batches; every kernel of a batch returns at once when the
// device-side stop test has fired, so the host synchronises once per batch, not per iteration.
void Handle::cg_solve(double tol, int max_iters, const double *x0, const double *rhs_)
{
    // r = rhs_ - KKT*x0 ; p = r ; rn = r.r   (x0/rhs_ default to sol/rhs; the fused path passes d/rhs')
    if (!x0) x0 = sol.p;
    if (!rhs_) rhs_ = rhs.p;
    FOS_LAUNCH(this, k_cg_begin, 1, 1, 0, d_ctrl.p, tol, max_iters);
    MVView V = kkt_pass(x0, nullptr);
    if (L.form == 0)
        FOS_LAUNCH(this, k2_kkt_hsde<K2_RESID>, vgrid(L.LP), VBLOCK, 0, L, V, x0, d_c.p, d_b.p, nullptr, rhs_, r.p,
                   p.p, d_ctrl.p, rb, 0);
    else if (L.form == 1)
        FOS_LAUNCH(this, k2_kkt_plain<K2_RESID>, vgrid(L.NP), VBLOCK, 0, L, V, x0, nullptr, rhs_, r.p, p.p, d_ctrl.p,
                   rb, 0);
    else
        FOS_LAUNCH(this, k2_spd<K2_RESID>, vgrid(L.NP), VBLOCK, 0, L, V, x0, nullptr, rhs_, r.p, p.p, d_ctrl.p, rb, 0);
    int64_t enq = 0;
    int batch = cg_batch > 0 ? cg_batch : (int)std::max<int64_t>(1, cgiter);  // adaptive: last solve's count
    for (;;) {
        batch = (int)std::min<int64_t>(batch, (int64_t)max_iters - enq);
        if (batch < 1) batch = 1;
        for (int k = 0; k < batch; k++) cg_enqueue_iteration();
        enq += batch;
        sync_ctrl();
        if (h_ctrl->done) break;
        FOS_REQUIRE(enq < (int64_t)max_iters + 4, "CG did not terminate (internal error)");
        batch = cg_batch > 0 ? cg_batch : std::max(1, batch / 2);
    }
    cgiter = h_ctrl->iter;
    stats.total_cg += cgiter;
    stats.total_passes += cgiter;  // one pass over A per executed CG iteration
    if (h_ctrl->warn_maxit) warn_maxit = true;
}

// prox!(y, S::AffinePlusLinear, x) (affinepluslinear.jl:83-126).  Result: sol (= xinit).
void Handle::s1_prox(const double *xin)
{
    if (direct) {  // HSDE.jl:10-15
        s1_prox_direct(xin);
        return;
    }
    const double an_ = (double)(L.form == 0 ? (L.n + L.m + 1) : L.n);
    if (fuse_rhs) {
        // one pass fewer per projection: see k_fuse_prep.  Same mathematics, different association of
        // the sums than affinepluslinear.jl:94-95 + conjugategradients.jl:32-33 (set "fuse_rhs" = 0 to
        // build rhs exactly in the reference's order).
        if (firstrun) {  // :101-104
            FOS_CUDA(cudaMemcpyAsync(sol.p, xin, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));
            firstrun = false;
        }
        FOS_LAUNCH(this, k_fuse_prep, vgrid(L.NP), VBLOCK, 0, L, sol.p, xin, beta, d_q.p, d_bhat.p, Ap.p, rhs.p);
        const double floor2 = an_ * 2.220446049250313e-16;
        double tol2 = floor2;
        if (decreasing) tol2 = std::max(std::pow(0.2, std::sqrt((double)s1_calls)), floor2);  // :108-112
        s1_calls += 1;                                                                        // :114
        cg_solve(tol2, 1000, Ap.p, L.form == 0 ? xin : rhs.p);
        return;
    }
    if (L.form == 0) {
        const double *X[1] = {xin + L.LP};
        const double *W[1] = {xin + L.LP + L.n_pad};
        MVView V = A.run(1, X, W, nullptr, stream);  // :94  (Q'x2 = -(Q x2))
        FOS_LAUNCH(this, k2_rhs_hsde, vgrid(L.LP), VBLOCK, 0, L, V, xin, d_c.p, d_b.p, rhs.p, rb);  // :95
    } else {
        const double *X[1] = {xin};
        const double *W[1] = {xin + L.n_pad};
        MVView V = A.run(1, X, W, nullptr, stream);
        FOS_LAUNCH(this, k2_rhs_plain, vgrid(L.n), VBLOCK, 0, L, V, xin, d_q.p, beta, rhs.p);
        FOS_CUDA(cudaMemcpyAsync(rhs.p + L.n_pad, d_bhat.p, (size_t)L.m_pad * 8, cudaMemcpyDeviceToDevice, stream));
    }
    if (firstrun) {  // :101-104
        FOS_CUDA(cudaMemcpyAsync(sol.p, xin, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));
        firstrun = false;
    }
    const double an = (double)(L.form == 0 ? (L.n + L.m + 1) : L.n);
    const double floor_ = an * 2.220446049250313e-16;
    double tol = floor_;
    if (decreasing) tol = std::max(std::pow(0.2, std::sqrt((double)s1_calls)), floor_);  // :108-112
    s1_calls += 1;                                                                       // :114
    cg_solve(tol, 1000);                                                                 // :115-118
}

// y = sol with y2 .*= beta (affinepluslinear.jl:124)
void Handle::sol_scaled_to(double *dst)
{
    FOS_LAUNCH(this, k_copy_scaled, vgrid(L.NP), VBLOCK, 0, L.NP, dst, sol.p, L.n_pad, L.form == 1 ? beta : 1.0);
}

void Handle::check(const double *z, int64_t i, bool override_)
{
    const bool due = (i % cur_checki == 0) || override_;
    if (L.form == 0) {
        if (!due) {  // HSDEStatus.jl:66-69
            checked = false;
            return;
        }
        const double *X[1] = {z};
        const double *W[1] = {z + L.n_pad};
        MVView V = A.run(1, X, W, nullptr, stream);
        FOS_LAUNCH(this, k6_check_hsde, vgrid(L.LP), VBLOCK, 0, L, V, z, d_c.p, d_b.p, nb, ncn, cur_eps, i,
                   (int)cgiter, d_ctrl.p, d_recs.p, rec_cap, rb);
    } else {
        FOS_LAUNCH(this, k6_check_feas, vgrid(L.NP), VBLOCK, 0, L.NP, z, prev.p, due ? 1 : 0, cur_eps, i, d_ctrl.p,
                   d_recs.p, rec_cap, rb);
        if (!due) {
            checked = false;
            return;
        }
    }
    sync_ctrl();
    status = h_ctrl->status;
    checked = true;
}

// =======================================================================================
// algorithms
// =======================================================================================
void Handle::set_algorithm(int alg_, double a, double a1, double a2, double bA, int64_t ip)
{
    FOS_REQUIRE(alg_ >= FOS_ALG_GAP && alg_ <= FOS_ALG_GAPP, "unknown algorithm code");
    FOS_REQUIRE(ip >= 1, "iproj must be >= 1");
    drop_graphs();  // alpha, alpha1, alpha2, beta are arguments of the captured kernels
    alg = alg_;
    alpha = a;
    alpha1 = a1;
    alpha2 = a2;
    betaA = bA;
    iproj = ip;
    fista_t = 1.0;  // fista.jl:24
    if (batch && batch->loaded) batch->set_algorithm();
    if (loaded) {
        const size_t bytes = (size_t)L.NP * 8;
        FOS_CUDA(cudaMemsetAsync(fy.p, 0, bytes, stream));
        FOS_CUDA(cudaMemsetAsync(fxold.p, 0, bytes, stream));
        FOS_CUDA(cudaMemsetAsync(dp.p, 0, bytes, stream));  // dykstra.jl:22
        FOS_CUDA(cudaMemsetAsync(dq.p, 0, bytes, stream));
        sync_ctrl();
        h_ctrl->alpha12 = 2.0;  // gapa.jl:29
        FOS_CUDA(cudaMemcpyAsync(d_ctrl.p, h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, stream));
        FOS_CUDA(cudaStreamSynchronize(stream));
    }
}

void Handle::step(int64_t i)
{
    const int g = vgrid(L.NP);
    const int64_t p2off = L.n_pad;
    const double p2scale = L.form == 1 ? beta : 1.0;
    EpiArgs E{};
    E.tmp2 = tmp2.p;
    E.x = x.p;
    E.betaA = betaA;
    switch (alg) {
    case FOS_ALG_GAP:
    case FOS_ALG_GAPA: {
        const bool ada = alg == FOS_ALG_GAPA;
        if (lsinterval > 0 && i % lsinterval == 0) {
            // step(::LineSearchWrapper, ...) on a line-search iteration (wrappers/linesearch.jl:42-72)
            double *t1s = w3.p, *res = w2.p, *t3 = w1.p;
            FOS_CUDA(cudaMemcpyAsync(t1s, x.p, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));  // :43
            s1_prox(x.p);                                                                           // S1! :47
            FOS_LAUNCH(this, k_relax, g, VBLOCK, 0, L.NP, tmp1.p, alpha1, sol.p, 1.0 - alpha1, x.p, p2off, p2scale,
                       d_ctrl.p, ada ? 1 : 0);
            E.a2 = alpha2;
            E.om_a2 = 1.0 - alpha2;
            E.use_a12 = ada ? 1 : 0;
            cone_project(cones, tmp1.p, proj.p, EPI_GAPP_PROJ, E);  // S2! :48: x = a2*P2(tmp2) + (1-a2)*tmp2
            check(proj.p, i, false);                                // checkstatus inside S2! (gap.jl:56)
            FOS_LAUNCH(this, k_sub, g, VBLOCK, 0, L.NP, res, x.p, t1s);  // :51
            FOS_LAUNCH(this, k_ls_begin, 1, 1, 0, d_ctrl.p, 1.0);
            double at = 0.1;
            for (int k = 0; k <= 30; k++) {  // :58-70, NoStatus
                at = at * 1.8;               // :59
                FOS_LAUNCH(this, k_add_scaled, g, VBLOCK, 0, L.NP, x.p, t1s, at, res, d_ctrl.p, 0);  // :60
                s1_prox(x.p);
                FOS_LAUNCH(this, k_relax, g, VBLOCK, 0, L.NP, tmp1.p, alpha1, sol.p, 1.0 - alpha1, x.p, p2off, p2scale,
                           d_ctrl.p, ada ? 1 : 0);
                E.ls_alpha = at;
                cone_project(cones, tmp1.p, t3, EPI_LSW, E);
            }
            FOS_LAUNCH(this, k_add_scaled, g, VBLOCK, 0, L.NP, x.p, t1s, 0.0, res, d_ctrl.p, 1);  // :72
            break;
        }
        s1_prox(x.p);  // gap.jl:45
        FOS_LAUNCH(this, k_relax, g, VBLOCK, 0, L.NP, tmp1.p, alpha1, sol.p, 1.0 - alpha1, x.p, p2off, p2scale,
                   d_ctrl.p, ada ? 1 : 0);  // :48
        E.a2 = alpha2;
        E.om_a2 = 1.0 - alpha2;
        E.a = alpha;
        E.om_a = 1.0 - alpha;
        cone_project(cones, tmp1.p, proj.p, ada ? EPI_GAPA : EPI_GAP, E);  // :55,58,78
        check(proj.p, i, false);                                           // :56
        break;
    }
    case FOS_ALG_FISTA: {
        if (i == 1) FOS_CUDA(cudaMemcpyAsync(fy.p, x.p, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));  // :31-33
        s1_prox(fy.p);  // :35
        FOS_LAUNCH(this, k_relax, g, VBLOCK, 0, L.NP, tmp1.p, alpha, sol.p, 1.0 - alpha, fy.p, p2off, p2scale,
                   d_ctrl.p, 0);  // :37
        const double told = fista_t;
        fista_t = (1.0 + std::sqrt(1.0 + 4.0 * told * told)) / 2.0;  // :45
        E.coef = (told - 1.0) / fista_t;                             // :46
        E.aux1 = fxold.p;
        E.aux2 = fy.p;
        cone_project(cones, tmp1.p, proj.p, EPI_FISTA, E);  // :39-41,46
        check(proj.p, i, false);
        break;
    }
    case FOS_ALG_DYKSTRA: {
        FOS_LAUNCH(this, k_add_scaled, g, VBLOCK, 0, L.NP, w1.p, x.p, 1.0, dp.p, d_ctrl.p, 0);  // x + p
        s1_prox(w1.p);                                                                          // :29
        sol_scaled_to(dy.p);
        FOS_LAUNCH(this, k_dykstra_mid, g, VBLOCK, 0, L.NP, w1.p, dy.p, dp.p, dq.p, w2.p);  // :31, y + q
        E.aux1 = dq.p;
        cone_project(cones, w2.p, proj.p, EPI_DYKSTRA, E);  // :32,35
        check(proj.p, i, false);                            // :33
        break;
    }
    case FOS_ALG_GAPP: {
        s1_prox(x.p);  // gapproj.jl:33
        E.a2 = alpha2;
        E.om_a2 = 1.0 - alpha2;
        E.a = alpha;
        E.om_a = 1.0 - alpha;
        if (i % iproj == 0) {  // :34
            sol_scaled_to(tmp1.p);
            cone_project(cones, tmp1.p, proj.p, EPI_NONE, E);  // :39
            s1_prox(proj.p);                                   // :40
            sol_scaled_to(w3.p);
            FOS_LAUNCH(this, k_sub, g, VBLOCK, 0, L.NP, w3.p, w3.p, tmp1.p);  // :41 res
            FOS_LAUNCH(this, k_ls_begin, 1, 1, 0, d_ctrl.p, -1.0);
            for (int k = 0; k <= 20; k++) {  // :46-56
                const double at = std::ldexp(1.0, k);
                FOS_LAUNCH(this, k_add_scaled, g, VBLOCK, 0, L.NP, w1.p, tmp1.p, at, w3.p, d_ctrl.p, 0);
                E.ls_alpha = at;
                cone_project(cones, w1.p, w2.p, EPI_LS, E);
            }
            FOS_LAUNCH(this, k_add_scaled, g, VBLOCK, 0, L.NP, tmp1.p, tmp1.p, 0.0, w3.p, d_ctrl.p, 1);  // :58
            cone_project(cones, tmp1.p, proj.p, EPI_GAPP_PROJ, E);                                      // :59,61,62
            check(proj.p, i, false);                                                                    // :60
        } else {
            FOS_LAUNCH(this, k_relax, g, VBLOCK, 0, L.NP, tmp1.p, alpha1, sol.p, 1.0 - alpha1, x.p, p2off, p2scale,
                       d_ctrl.p, 0);                          // :64
            cone_project(cones, tmp1.p, proj.p, EPI_GAP, E);  // :66,68,70
            check(proj.p, i, false);                          // :67
        }
        break;
    }
    default: throw Error(FOS_ERR_INVALID, "unknown algorithm");
    }
}

// getsol (gap.jl:82-87 and the identical methods of the other algorithms): P2(P1(x)).
// Result in proj (and tmp1/tmp2 hold what the reference leaves there).
void Handle::getsol()
{
    s1_prox(x.p);
    sol_scaled_to(tmp1.p);
    EpiArgs E{};
    cone_project(cones, tmp1.p, proj.p, EPI_NONE, E);
    if (alg == FOS_ALG_DYKSTRA)
        FOS_CUDA(cudaMemcpyAsync(dy.p, proj.p, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));
    else
        FOS_CUDA(cudaMemcpyAsync(tmp2.p, proj.p, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));
}

// fresh status object, as status_generator builds one per solve! (HSDE.jl:26-27, Feasibility.jl:78-79)
void Handle::begin_solve()
{
    status = FOS_STATUS_CONTINUE;
    checked = false;
    // a fresh status object counts from i = 0 (a forced final check after max_iters = 0 records i = 0), and the
    // "CG reached max iterations" warning is raised per occurrence (conjugategradients.jl:53), not once per model
    cur_i = 0;
    warn_maxit = false;
    FOS_CUDA(cudaMemsetAsync(&d_ctrl.p->warn_maxit, 0, sizeof(int32_t), stream));
    if (L.form == 1) {
        std::vector<double> nanv((size_t)L.NP, std::nan(""));
        FOS_CUDA(cudaMemcpy(prev.p, nanv.data(), (size_t)L.NP * 8, cudaMemcpyHostToDevice));
        FOS_SYNC_LEGACY();
    }
}

int64_t Handle::run(int64_t i_start, int64_t n_iters, int64_t checki, double eps, double *records,
                    int64_t rec_cap_host, int64_t *n_rec, double *trace)
{
    require_loaded();
    FOS_REQUIRE(checki >= 1, "checki must be >= 1");
    FOS_REQUIRE(n_iters >= 0, "n_iters must be >= 0");
    cur_checki = checki;
    cur_eps = eps;
    ensure_recs((int)std::min<int64_t>(n_iters / checki + 2, 1 << 20));
    sync_ctrl();
    h_ctrl->nrec = 0;
    h_ctrl->status = status;
    h_ctrl->s1_calls_dev = s1_calls;
    h_ctrl->cur_i = i_start - 1;
    h_ctrl->total_cg_dev = 0;
    h_ctrl->fista_t = fista_t;
    FOS_CUDA(cudaMemcpyAsync(d_ctrl.p, h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, stream));
    int64_t done = 0;
    bool synced = false;
    if (!trace && graph_ok()) {
        done = run_graph(i_start, n_iters);
        synced = true;  // run_graph ends with a synchronisation that fetched the control block
    } else {
        for (int64_t i = i_start; i < i_start + n_iters; i++) {
            cur_i = i;  // solverwrapper.jl:24
            step(i);    // :25
            if (trace) unpack_to_host(x.p, trace + done * N);
            done++;
            if (status != FOS_STATUS_CONTINUE) break;  // :26-28
        }
    }
    if (!synced) sync_ctrl();
    const int64_t nrec = h_ctrl->nrec;
    if (n_rec) *n_rec = nrec;
    const int64_t ncopy = std::min<int64_t>(std::min<int64_t>(nrec, rec_cap_host), rec_cap);
    if (records && ncopy > 0)
        FOS_CUDA(cudaMemcpy(records, d_recs.p, (size_t)ncopy * FOS_REC_LEN * 8, cudaMemcpyDeviceToHost));
    return done;
}

// tail of iterate() (solverwrapper.jl:31-34)
void Handle::finish(double *guess, double *record, int64_t *n_rec)
{
    require_loaded();
    sync_ctrl();
    h_ctrl->nrec = 0;
    FOS_CUDA(cudaMemcpyAsync(d_ctrl.p, h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, stream));
    getsol();
    int64_t nrec = 0;
    if (!checked) {
        check(proj.p, cur_i, true);
        nrec = 1;
        if (record) FOS_CUDA(cudaMemcpy(record, d_recs.p, FOS_REC_LEN * 8, cudaMemcpyDeviceToHost));
    }
    if (n_rec) *n_rec = nrec;
    if (guess) unpack_to_host(proj.p, guess);
    else FOS_CUDA(cudaStreamSynchronize(stream));
}


// =======================================================================================
// graph path: one CUDA graph per outer iteration, the CG loop as a WHILE node
// =======================================================================================
bool Handle::graph_ok() const
{
    return use_graphs && loaded && L.form == 0 &&
           (alg == FOS_ALG_GAP || alg == FOS_ALG_GAPA || alg == FOS_ALG_FISTA || alg == FOS_ALG_DYKSTRA) && lsinterval == 0 &&
           !direct && fuse_rhs && fuse_tail && A.kind == 1 && A.impl == 0 && cones.fusable && !A.profile &&
           (A.nranks == 1 || A.p2p_on) && L.NP > 0;
}

void Handle::drop_graphs()
{
    for (IterGraph &g : graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
        g.exec = nullptr;
        g.graph = nullptr;
        g.fixed_launches = 0;
    }
}

void Handle::build_iter_graph(bool with_check)
{
    IterGraph &G = graphs[with_check ? 1 : 0];
    const bool p2p = A.nranks > 1 && A.p2p_on;
    const bool ada = alg == FOS_ALG_GAPA;
    {   // everything that may not happen while the stream is capturing
        int &occ = tail_occ[p2p ? 1 : 0];
        if (occ == 0) {
            const void *fn = p2p ? (const void *)k_cg_tail_hsde<true> : (const void *)k_cg_tail_hsde<false>;
            FOS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, VBLOCK, 0));
            FOS_REQUIRE(occ >= 1, "the fused CG tail does not fit on an SM");
        }
        FOS_CUDA(cudaStreamSynchronize(stream));
    }
    const Stats saved = stats;  // launches made while capturing are recorded, not executed
    cudaStream_t main_stream = stream;
    cudaGraph_t graph = nullptr;
    capturing = true;
    try {
        FOS_CUDA(cudaStreamBeginCapture(main_stream, cudaStreamCaptureModeThreadLocal));
        const int64_t l0 = stats.launches;
        // ---- the point S1 projects: x (GAP family, gap.jl:45), y (FISTA, fista.jl:35), x + p (Dykstra, dykstra.jl:29) ----
        const double *xin = x.p;
        if (alg == FOS_ALG_FISTA) xin = fy.p;
        if (alg == FOS_ALG_DYKSTRA) {
            FOS_LAUNCH(this, k_add_scaled, vgrid(L.NP), VBLOCK, 0, L.NP, w1.p, x.p, 1.0, dp.p, d_ctrl.p, 0);
            xin = w1.p;
        }
        // ---- AffinePlusLinear.prox!, first part: right-hand side folded into the initial residual ----
        FOS_LAUNCH(this, k_iter_begin, vgrid(L.NP), VBLOCK, 0, L, sol.p, xin, Ap.p, d_ctrl.p, tol_table.p, TOL_TABLE_N,
                   1000, alg == FOS_ALG_FISTA ? 1 : 0);
        MVView V = kkt_pass(Ap.p, nullptr, /*defer_exchange=*/p2p);
        if (p2p) {
            const int grid = (int)std::min<int64_t>((L.LP + VBLOCK - 1) / VBLOCK, 2 * (int64_t)num_sms);
            FOS_LAUNCH(this, k2_resid_hsde_p2p, grid, VBLOCK, 0, L, V, A.p2p, Ap.p, d_c.p, d_b.p, xin, r.p, p.p,
                       d_ctrl.p, rb);
        } else {
            FOS_LAUNCH(this, k2_kkt_hsde<K2_RESID>, vgrid(L.LP), VBLOCK, 0, L, V, Ap.p, d_c.p, d_b.p, nullptr, xin, r.p,
                       p.p, d_ctrl.p, rb, 0);
        }
        // ---- conjugategradient!'s loop (conjugategradients.jl:37-52) as a WHILE node; always >= 1 trip ----
        cudaStreamCaptureStatus cst;
        unsigned long long cid = 0;
        cudaGraph_t cgraph = nullptr;
        const cudaGraphNode_t *deps = nullptr;
        size_t ndeps = 0;
        FOS_CUDA(cudaStreamGetCaptureInfo(main_stream, &cst, &cid, &cgraph, &deps, &ndeps));
        cudaGraphConditionalHandle cond;
        FOS_CUDA(cudaGraphConditionalHandleCreate(&cond, cgraph, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t wnode;
        FOS_CUDA(cudaGraphAddNode(&wnode, cgraph, deps, ndeps, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        FOS_CUDA(cudaStreamBeginCaptureToGraph(stream2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        stream = stream2;
        const int64_t lb = stats.launches;
        try {
            cg_enqueue_iteration((unsigned long long)cond);
        } catch (...) {
            stream = main_stream;
            cudaGraph_t dummy = nullptr;
            cudaStreamEndCapture(stream2, &dummy);
            throw;
        }
        stream = main_stream;
        const int64_t body_launches = stats.launches - lb;
        FOS_CUDA(cudaStreamEndCapture(stream2, nullptr));
        FOS_CUDA(cudaStreamUpdateCaptureDependencies(main_stream, &wnode, 1, cudaStreamSetCaptureDependencies));
        // ---- S1 relaxation + S2 = DualConeProduct + the algorithm's epilogue: one kernel (two with SOC cones) ----
        EpiArgs E{};
        E.tmp2 = tmp2.p;
        E.x = x.p;
        E.betaA = betaA;
        E.a2 = alpha2;
        E.om_a2 = 1.0 - alpha2;
        E.a = alpha;
        E.om_a = 1.0 - alpha;
        RelaxArgs R{};
        R.X = sol.p;
        R.tmp1 = tmp1.p;
        ConeSet &K = cones;
        const int gK = vgrid(K.NP);
        if (alg == FOS_ALG_DYKSTRA) {
            // p = (x + p) - y ; y + q -> P2 ; x = P2(y + q) ; q = (y + q) - x   (dykstra.jl:31-35), y = sol
            FOS_LAUNCH(this, k_dykstra_mid, gK, VBLOCK, 0, L.NP, w1.p, sol.p, dp.p, dq.p, w2.p);
            E.aux1 = dq.p;
            if (K.nsoc > 0)
                FOS_LAUNCH(this, k4_soc_norms<false>, K.nchunks, VBLOCK, 0, w2.p, K.soc.p, K.nsoc, K.chunk_cone.p,
                           K.chunk_sum.p, K.soc_scale.p, K.counter.p, R, d_ctrl.p);
            FOS_LAUNCH(this, (k4_cone_apply<EPI_DYKSTRA, false>), gK, VBLOCK, 0, K.NP, w2.p, proj.p, K.ops.p, K.cone_of.p,
                       K.soc_scale.p, K.soc.p, K.box.p, E, d_ctrl.p, rb, R);
        } else {
            const bool fis = alg == FOS_ALG_FISTA;
            R.Y = fis ? fy.p : x.p;                 // fista.jl:37 relaxes against y with alpha; gap.jl:48 against x with alpha1
            R.a = fis ? alpha : alpha1;
            R.b = 1.0 - R.a;
            R.use_a12 = ada ? 1 : 0;
            if (fis) {
                E.aux1 = fxold.p;
                E.aux2 = fy.p;
                E.coef_from_ctrl = 1;
            }
            if (K.nsoc > 0)
                FOS_LAUNCH(this, k4_soc_norms<true>, K.nchunks, VBLOCK, 0, nullptr, K.soc.p, K.nsoc, K.chunk_cone.p,
                           K.chunk_sum.p, K.soc_scale.p, K.counter.p, R, d_ctrl.p);
            if (fis)
                FOS_LAUNCH(this, (k4_cone_apply<EPI_FISTA, true>), gK, VBLOCK, 0, K.NP, nullptr, proj.p, K.ops.p,
                           K.cone_of.p, K.soc_scale.p, K.soc.p, K.box.p, E, d_ctrl.p, rb, R);
            else if (ada)
                FOS_LAUNCH(this, (k4_cone_apply<EPI_GAPA, true>), gK, VBLOCK, 0, K.NP, nullptr, proj.p, K.ops.p,
                           K.cone_of.p, K.soc_scale.p, K.soc.p, K.box.p, E, d_ctrl.p, rb, R);
            else
                FOS_LAUNCH(this, (k4_cone_apply<EPI_GAP, true>), gK, VBLOCK, 0, K.NP, nullptr, proj.p, K.ops.p,
                           K.cone_of.p, K.soc_scale.p, K.soc.p, K.box.p, E, d_ctrl.p, rb, R);
        }
        // ---- checkstatus on the unrelaxed projection (gap.jl:56; HSDEStatus.jl:27-71) ----
        if (with_check) {
            const double *X1[1] = {proj.p};
            const double *W1[1] = {proj.p + L.n_pad};
            MVView Vc = A.run(1, X1, W1, nullptr, main_stream);
            FOS_LAUNCH(this, k6_check_hsde, vgrid(L.LP), VBLOCK, 0, L, Vc, proj.p, d_c.p, d_b.p, nb, ncn, cur_eps,
                       (int64_t)-1, -1, d_ctrl.p, d_recs.p, rec_cap, rb);
        }
        G.fixed_launches = (int)(stats.launches - l0 - body_launches);
        FOS_CUDA(cudaStreamEndCapture(main_stream, &graph));
    } catch (...) {
        stream = main_stream;
        capturing = false;
        stats = saved;
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(main_stream, &cst) == cudaSuccess && cst != cudaStreamCaptureStatusNone) {
            cudaGraph_t dummy = nullptr;
            cudaStreamEndCapture(main_stream, &dummy);
            if (dummy) cudaGraphDestroy(dummy);
        }
        cudaGetLastError();
        throw;
    }
    capturing = false;
    stats = saved;
    G.graph = graph;
    FOS_CUDA(cudaGraphInstantiate(&G.exec, graph, 0));
    if (with_check) graph_eps = cur_eps;
}

// iterations i_start .. i_start+n_iters-1 of solverwrapper.jl:23-29; the host synchronises on check iterations only
int64_t Handle::run_graph(int64_t i_start, int64_t n_iters)
{
    if (graphs[1].exec && graph_eps != cur_eps) {  // eps is an argument of the captured check kernel
        cudaGraphExecDestroy(graphs[1].exec);
        cudaGraphDestroy(graphs[1].graph);
        graphs[1] = IterGraph();
    }
    int64_t done = 0, since_iters = 0, since_checks = 0, absorbed = 0;
    auto absorb = [&]() {
        sync_ctrl();
        const int64_t delta = h_ctrl->total_cg_dev - absorbed;
        absorbed += delta;
        stats.total_cg += delta;
        stats.total_passes += delta + since_iters + since_checks;  // k CG passes + the initial residual (+ the check)
        stats.launches += (since_iters - since_checks) * graphs[0].fixed_launches +
                          since_checks * graphs[1].fixed_launches + 2 * delta;
        cgiter = h_ctrl->iter;
        if (h_ctrl->warn_maxit) warn_maxit = true;
        if (alg == FOS_ALG_FISTA) fista_t = h_ctrl->fista_t;  // the device advanced t (k_iter_begin)
        since_iters = 0;
        since_checks = 0;
    };
    for (int64_t i = i_start; i < i_start + n_iters; i++) {
        if (alg == FOS_ALG_FISTA && i == 1)  // fista.jl:31-33
            FOS_CUDA(cudaMemcpyAsync(fy.p, x.p, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));
        if (firstrun) {  // affinepluslinear.jl:101-104: the first point S1 ever projects is its own warm start
            const double *first = alg == FOS_ALG_FISTA ? fy.p : x.p;  // Dykstra: x + p with p = 0 (bitwise x) on a first run
            if (alg == FOS_ALG_DYKSTRA)
                FOS_LAUNCH(this, k_add_scaled, vgrid(L.NP), VBLOCK, 0, L.NP, sol.p, x.p, 1.0, dp.p, d_ctrl.p, 0);
            else
                FOS_CUDA(cudaMemcpyAsync(sol.p, first, (size_t)L.NP * 8, cudaMemcpyDeviceToDevice, stream));
            firstrun = false;
        }
        const bool chk = (i % cur_checki) == 0;
        IterGraph &G = graphs[chk ? 1 : 0];
        if (!G.exec) build_iter_graph(chk);
        FOS_CUDA(cudaGraphLaunch(G.exec, stream));
        cur_i = i;
        s1_calls += 1;
        done++;
        since_iters++;
        if (chk) {
            since_checks++;
            absorb();
            status = h_ctrl->status;
            checked = true;
            if (status != FOS_STATUS_CONTINUE) break;
        } else {
            checked = false;
        }
    }
    absorb();
    return done;
}

}  // namespace fos
