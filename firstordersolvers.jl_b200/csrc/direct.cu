// direct.cu -- direct = true: S1 = IndAffine([Q -I], 0), the EXACT projection onto {(u, v) : Q u = v}
// (problemforms/HSDE/HSDE.jl:10-15; the default of GAPP, solvers/gapproj.jl:14).
//
// ProximalOperators factorises the normal equations once and solves them every iteration.  Here the
// one-time work runs on the GPU at load time and the per-iteration work is three streaming passes:
//   setup    Qc   = Q (padded coordinates, column j contiguous), one fused mat-vec per unit vector
//            G    = I + Q Q'             64x64 register-tiled Gram kernel
//            W    = G^-1                 in-place Gauss-Jordan (G is SPD with eigenvalues >= 1: no pivoting)
//   prox     w    = Q u - v              K1 pass over A  (k2_q_hsde)
//            t    = W w                  K1 pass over W  (the same TMA kernel, on the l x l matrix)
//            y    = [u - Q't ; v + t]    K1 pass over A
// i.e. 2 passes over A + 1 pass over W per projection, against k+1 passes over A for the CG solve;
// meant for small / medium l (W has l^2 entries).  No CG, no warm start, no tolerance schedule.
#include <algorithm>

#include "solver.cuh"

namespace fos {

namespace {

constexpr int GT = 64;  // Gram tile
constexpr int GJ = 32;  // columns of Qc per shared-memory chunk

// G = I + Qc' Qc over padded coordinates: G[i][k] = delta_ik + sum_j Qc[j*LP + i] * Qc[j*LP + k]
__global__ void __launch_bounds__(256) k_gram_plus_identity(const double *__restrict__ Qc, int64_t LP, double *__restrict__ G)
{
    __shared__ __align__(16) double Wi[GJ * GT], Wk[GJ * GT];
    const int I = blockIdx.y, K = blockIdx.x;
    const int ti = threadIdx.x & 15, tk = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) acc[x][y] = 0.0;
    for (int64_t j0 = 0; j0 < LP; j0 += GJ) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < GJ * GT; idx += 256) {
            const int jj = idx / GT, r = idx - jj * GT;
            const int64_t j = j0 + jj;
            const int64_t gi = (int64_t)I * GT + r, gk = (int64_t)K * GT + r;
            Wi[idx] = (j < LP && gi < LP) ? Qc[j * LP + gi] : 0.0;
            Wk[idx] = (j < LP && gk < LP) ? Qc[j * LP + gk] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < GJ; jj++) {
            const double2 i01 = *reinterpret_cast<const double2 *>(Wi + jj * GT + 4 * ti);
            const double2 i23 = *reinterpret_cast<const double2 *>(Wi + jj * GT + 4 * ti + 2);
            const double2 k01 = *reinterpret_cast<const double2 *>(Wk + jj * GT + 4 * tk);
            const double2 k23 = *reinterpret_cast<const double2 *>(Wk + jj * GT + 4 * tk + 2);
            const double wi[4] = {i01.x, i01.y, i23.x, i23.y};
            const double wk[4] = {k01.x, k01.y, k23.x, k23.y};
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y] = fma(wi[x], wk[y], acc[x][y]);
        }
    }
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const int64_t i = (int64_t)I * GT + 4 * ti + x, k = (int64_t)K * GT + 4 * tk + y;
            if (i < LP && k < LP) G[i * LP + k] = acc[x][y] + (i == k ? 1.0 : 0.0);
        }
}

// In-place Gauss-Jordan inversion, pivot k:
//   pivot kernel : colk[i] = G[i][k] (i != k), G[i][k] = 0 ; row k /= p with G[k][k] = 1/p
//   update kernel: G[i][j] -= colk[i] * G[k][j] for every i != k
__global__ void k_gj_begin(double *__restrict__ G, int64_t LP, int64_t k, double *__restrict__ piv)
{
    piv[0] = G[k * LP + k];
    G[k * LP + k] = 1.0;  // so that the row scaling below leaves 1/p on the diagonal
}
__global__ void __launch_bounds__(256) k_gj_pivot(double *__restrict__ G, int64_t LP, int64_t k, const double *__restrict__ piv,
                                                  double *__restrict__ colk)
{
    const double p = piv[0];
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < LP; e += (int64_t)gridDim.x * 256) {
        if (e != k) {
            colk[e] = G[e * LP + k];
            G[e * LP + k] = 0.0;
        } else {
            colk[e] = 0.0;
        }
        G[k * LP + e] = G[k * LP + e] / p;
    }
}
__global__ void __launch_bounds__(256) k_gj_update(double *__restrict__ G, int64_t LP, int64_t k, const double *__restrict__ colk)
{
    const int64_t i = blockIdx.y;
    if (i == k) return;
    const double f = colk[i];
    if (f == 0.0) return;
    const double *rowk = G + k * LP;
    double *rowi = G + i * LP;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < LP; j += (int64_t)gridDim.x * 256)
        rowi[j] = fma(-f, rowk[j], rowi[j]);
}

__global__ void k_set_entry(double *v, int64_t idx, double val) { v[idx] = val; }

// w = Qu - v  (half-vectors of LP entries)
__global__ void __launch_bounds__(VBLOCK) k_direct_w(int64_t LP, const double *__restrict__ Qu, const double *__restrict__ v,
                                                     double *__restrict__ w)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < LP; e += (int64_t)gridDim.x * VBLOCK)
        w[e] = sub_(Qu[e], v[e]);
}
// t = W w from the pass over W
__global__ void __launch_bounds__(VBLOCK) k_direct_fold(int64_t LP, MVView V, double *__restrict__ t)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < LP; e += (int64_t)gridDim.x * VBLOCK)
        t[e] = mv_ax(V, 0, e);
}
// y = [u - Q't ; v + t]
__global__ void __launch_bounds__(VBLOCK) k_direct_out(int64_t LP, const double *__restrict__ x, const double *__restrict__ Qtt,
                                                       const double *__restrict__ t, double *__restrict__ y)
{
    for (int64_t e = (int64_t)blockIdx.x * VBLOCK + threadIdx.x; e < LP; e += (int64_t)gridDim.x * VBLOCK) {
        y[e] = sub_(x[e], Qtt[e]);
        y[LP + e] = add_(x[LP + e], t[e]);
    }
}

}  // namespace

void Handle::set_direct(bool on)
{
    require_loaded();
    FOS_REQUIRE(L.form == 0, "direct = true applies to the conic (HSDE) form; the Feasibility form takes its S1 from the user");
    FOS_REQUIRE(A.nranks == 1, "direct = true is not offered with row sharding");
    if (direct) {
        // s1_prox_direct borrows rhs and r as scratch; the CG path with "fuse_rhs" = 0 relies on rhs[LP..2LP) = b = 0
        // (HSDE.jl:22) and never rewrites it
        FOS_CUDA(cudaMemsetAsync(rhs.p, 0, (size_t)L.NP * 8, stream));
        FOS_CUDA(cudaMemsetAsync(r.p, 0, (size_t)L.NP * 8, stream));
        FOS_CUDA(cudaStreamSynchronize(stream));
    }
    direct = false;
    Wop.reset();
    Winv.release();
    if (!on) return;
    const int64_t LP = L.LP;
    if (LP > 16384)
        throw Error(FOS_ERR_UNSUPPORTED, "direct = true needs the dense " + std::to_string(LP) + " x " + std::to_string(LP) +
                                             " inverse of I + Q Q'; supported up to 16384 (use direct = false)");
    DevBuf<double> Qc, e, colk, piv;
    Qc.alloc((size_t)LP * LP);
    e.alloc((size_t)LP);
    colk.alloc((size_t)LP);
    piv.alloc(1);
    Winv.alloc((size_t)LP * LP, false);
    // columns of Q in padded coordinates (padding rows / columns stay zero)
    const int64_t offs[3] = {0, L.n_pad, L.n_pad + L.m_pad};
    const int64_t lens[3] = {L.n, L.m, 1};
    for (int s = 0; s < 3; s++)
        for (int64_t k = 0; k < lens[s]; k++) {
            const int64_t j = offs[s] + k;
            k_set_entry<<<1, 1, 0, stream>>>(e.p, j, 1.0);
            q_mul(e.p, Qc.p + (size_t)j * LP, false);
            k_set_entry<<<1, 1, 0, stream>>>(e.p, j, 0.0);
        }
    const unsigned nt = (unsigned)((LP + GT - 1) / GT);
    k_gram_plus_identity<<<dim3(nt, nt), 256, 0, stream>>>(Qc.p, LP, Winv.p);
    const unsigned gx = (unsigned)std::min<int64_t>((LP + 255) / 256, 8);
    for (int64_t k = 0; k < LP; k++) {
        k_gj_begin<<<1, 1, 0, stream>>>(Winv.p, LP, k, piv.p);
        k_gj_pivot<<<(unsigned)((LP + 255) / 256), 256, 0, stream>>>(Winv.p, LP, k, piv.p, colk.p);
        k_gj_update<<<dim3(gx, (unsigned)LP), 256, 0, stream>>>(Winv.p, LP, k, colk.p);
    }
    FOS_CUDA(cudaStreamSynchronize(stream));
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) throw Error(FOS_ERR_CUDA, std::string("direct set-up failed: ") + cudaGetErrorString(err));
    stats.launches += 3 * LP + 2 * (L.n + L.m + 1) + 1;
    Wop.reset(new MatOp());
    Wop->num_sms = num_sms;
    Wop->stats = &stats;
    Wop->impl = matvec_impl;
    Wop->init_dense(LP, LP, Winv.p, LP, FOS_MEM_DEVICE, 0, LP, grid_ctas, stream);
    direct = true;
}

// prox of IndAffine([Q -I], 0) at xin; result in sol
void Handle::s1_prox_direct(const double *xin)
{
    const int64_t LP = L.LP;
    const int g = vgrid(LP);
    double *Qu = rhs.p, *w = rhs.p + LP, *t = r.p, *Qtt = r.p + LP;
    q_mul(xin, Qu, false);
    FOS_LAUNCH(this, k_direct_w, g, VBLOCK, 0, LP, Qu, xin + LP, w);
    const double *X[1] = {w};
    const double *Wv[1] = {w};
    MVView V = Wop->run(1, X, Wv, nullptr, stream);
    FOS_LAUNCH(this, k_direct_fold, g, VBLOCK, 0, LP, V, t);
    q_mul(t, Qtt, true);
    FOS_LAUNCH(this, k_direct_out, g, VBLOCK, 0, LP, xin, Qtt, t, sol.p);
    cgiter = 0;
}

}  // namespace fos
