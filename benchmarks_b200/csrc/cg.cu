// cg.cu -- C ABI section 5: conjugate gradients with fused vector updates and device-side control.
//
// Mirrors dealii::SolverCG + ReductionControl as the reference drivers use them
// (CEED_bp/src/bp3.cc:266-285: PreconditionIdentity, ReductionControl(1e9, 1e-16, 1e-9);
//  bp5_kokkos/benchmark.cc:355-378: Jacobi, ReductionControl(100, 1e-15, 1e-8), NoConvergence
//  swallowed); loop restated from SURVEY.md appendix A9:
//     r = b (x0 = 0); res0 = |r|
//     it = 1..: rho_old = rho; rho = r.z (z = P^-1 r); p = z + (rho/rho_old) p (first: p = z)
//               v = A p; alpha = rho / (p.v); x += alpha p; r -= alpha v; res = |r|
//     stop when res <= abs_tol or res <= rel_tol * res0; last_step() = it
//
// B200 design: per iteration three kernels instead of deal.II's five vector sweeps + two host-synchronous
// reductions:   [scalar step]  ->  [p = z + beta p]  ->  [v = A p with p.v fused into the cell kernel]
//               ->  [x += alpha p; r -= alpha v; r.r and r.z fused]
// alpha/beta/convergence live in device memory; the host only polls a flag every `check_every`
// iterations, and iterations queued after convergence are no-ops, so x is exactly the iterate at
// which ReductionControl would have stopped.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <list>
#include <memory>
#include <mutex>
#include <vector>

#include "halo.h"
#include "operator.h"

namespace b200fe {

namespace {

struct CgScalars {      // device-resident
    double rho, rho_old;
    double acc[3];      // [0] p.Ap  [1] r.r  [2] r.z   (summed over ranks in place)
    double res0, res;
    double alpha_pending;  // x += alpha p of the iteration just finished is applied by the NEXT p-update pass (or the flush)
    int has_pending;
    int it;             // iterations started
    int done, converged, its;
};

// Vector-valued problems (BP2/BP4/BP6): component-blocked vectors [component][stride]; blockIdx.y = component,
// the (scalar-operator) inverse diagonal is shared by the components.
__global__ void cg_init_kernel(uint32_t n, size_t stride, const double *__restrict__ b, double *__restrict__ x, double *__restrict__ r,
                               const double *__restrict__ inv_diag, CgScalars *sc)
{
    b += blockIdx.y * stride; x += blockIdx.y * stride; r += blockIdx.y * stride;
    double rr = 0.0, rz = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double bi = b[i];
        x[i] = 0.0;
        r[i] = bi;
        rr = fma(bi, bi, rr);
        if (inv_diag) rz = fma(bi * inv_diag[i], bi, rz);
    }
    for (int o = 16; o > 0; o >>= 1) {
        rr += __shfl_xor_sync(0xffffffffu, rr, o);
        rz += __shfl_xor_sync(0xffffffffu, rz, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sc->acc[1], rr);
        if (inv_diag) atomicAdd(&sc->acc[2], rz);
    }
}

// ReductionControl check for the iterate just finished, then set up rho for the next iteration
__global__ void cg_scalar_kernel(CgScalars *sc, int jacobi, int max_it, double abs_tol, double rel_tol)
{
    if (sc->done) return;
    const double res = sqrt(fabs(sc->acc[1]));
    sc->res = res;
    if (sc->it == 0) sc->res0 = res;
    if (!(res == res) || isinf(res)) {  // breakdown (NaN/Inf): stop like deal.II's is_finite assertion, not converged
        sc->done = 1; sc->converged = 0; sc->its = sc->it;
        return;
    }
    if (res <= abs_tol || res <= rel_tol * sc->res0) {
        sc->done = 1; sc->converged = 1; sc->its = sc->it;
        return;
    }
    if (sc->it >= max_it) {
        sc->done = 1; sc->converged = 0; sc->its = sc->it;
        return;
    }
    sc->rho_old = sc->rho;
    sc->rho = jacobi ? sc->acc[2] : sc->acc[1];
    sc->acc[0] = sc->acc[1] = sc->acc[2] = 0.0;
    sc->it += 1;
}

// One pass over the vectors between two operator applications: p = z + beta p; the pending x += alpha p_old of the
// previous iteration (p_old is read here anyway: 8 B/DoF less than updating x together with r); v = 0 for the scatter of
// the coming apply (owned and ghost entries; replaces a separate memset launch).
__global__ void cg_update_p_kernel(uint32_t n, uint32_t n_local, size_t stride, const double *__restrict__ r, const double *__restrict__ inv_diag,
                                   double *__restrict__ p, double *__restrict__ x, double *__restrict__ v, const CgScalars *sc,
                                   const uint32_t *__restrict__ excl_mask)
{
    if (sc->done) return;
    r += blockIdx.y * stride; p += blockIdx.y * stride; x += blockIdx.y * stride; v += blockIdx.y * stride;
    const bool first = sc->it == 1;
    const double beta = first ? 0.0 : sc->rho / sc->rho_old;
    const bool pending = sc->has_pending != 0;
    const double alpha = sc->alpha_pending;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += gridDim.x * blockDim.x) {
        // (cell-interior DoFs are overwritten by the apply with plain stores: no zero-fill, Operator::d_excl_mask)
        if (excl_mask == nullptr || !((__ldg(excl_mask + (i >> 5)) >> (i & 31)) & 1u)) v[i] = 0.0;
        if (i < n) {
            const double z = inv_diag ? inv_diag[i] * r[i] : r[i];
            const double pi = first ? 0.0 : p[i];
            if (pending) x[i] = fma(alpha, pi, x[i]);
            p[i] = first ? z : fma(beta, pi, z);
        }
    }
}

// after the loop: the x update of the last iteration (p still holds its search direction: the p pass is a no-op once done)
__global__ void cg_flush_x_kernel(uint32_t n, size_t stride, const double *__restrict__ p, double *__restrict__ x, CgScalars *sc)
{
    if (!sc->has_pending) return;
    p += blockIdx.y * stride; x += blockIdx.y * stride;
    const double alpha = sc->alpha_pending;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = fma(alpha, p[i], x[i]);
}
__global__ void cg_clear_pending_kernel(CgScalars *sc) { sc->has_pending = 0; }

// r -= alpha v with r.r (and r.z) fused; x += alpha p is left pending for the next p pass (cg_update_p_kernel)
__global__ void cg_update_xr_kernel(uint32_t n, size_t stride, const double *__restrict__ v,
                                    const double *__restrict__ inv_diag, double *__restrict__ r, CgScalars *sc)
{
    if (sc->done) return;
    v += blockIdx.y * stride; r += blockIdx.y * stride;
    const double alpha = sc->rho / sc->acc[0];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {  // (every block reads rho and acc[0] only; separate fields)
        sc->alpha_pending = alpha;
        sc->has_pending = 1;
    }
    double rr = 0.0, rz = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double ri = fma(-alpha, v[i], r[i]);
        r[i] = ri;
        rr = fma(ri, ri, rr);
        if (inv_diag) rz = fma(ri * inv_diag[i], ri, rz);
    }
    for (int o = 16; o > 0; o >>= 1) {
        rr += __shfl_xor_sync(0xffffffffu, rr, o);
        rz += __shfl_xor_sync(0xffffffffu, rz, o);
    }
    __shared__ double s_rr[32], s_rz[32];
    if ((threadIdx.x & 31) == 0) { s_rr[threadIdx.x >> 5] = rr; s_rz[threadIdx.x >> 5] = rz; }
    __syncthreads();
    if (threadIdx.x < 32) {
        rr = threadIdx.x < (blockDim.x >> 5) ? s_rr[threadIdx.x] : 0.0;
        rz = threadIdx.x < (blockDim.x >> 5) ? s_rz[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            rr += __shfl_xor_sync(0xffffffffu, rr, o);
            rz += __shfl_xor_sync(0xffffffffu, rz, o);
        }
        if (threadIdx.x == 0) {
            // NB: acc[1]/acc[2] were zeroed by the scalar step; acc[0] is consumed above by every block
            atomicAdd(&sc->acc[1], rr);
            if (inv_diag) atomicAdd(&sc->acc[2], rz);
        }
    }
}

// ---- Chebyshev polynomial preconditioner on the Jacobi-scaled operator (deal.II PreconditionChebyshev) -------------------
// first term: d = c0 D^-1 r, z = d
__global__ void cheb_first_kernel(uint32_t n, double c0, const double *__restrict__ r, const double *__restrict__ inv_diag, double *__restrict__ d,
                                  double *__restrict__ z, const CgScalars *sc)
{
    if (sc != nullptr && sc->done) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double di = c0 * inv_diag[i] * r[i];
        d[i] = di;
        z[i] = di;
    }
}
// next term, t = A z given: d = a d + b D^-1 (r - t), z += d
__global__ void cheb_step_kernel(uint32_t n, double a, double b, const double *__restrict__ r, const double *__restrict__ t,
                                 const double *__restrict__ inv_diag, double *__restrict__ d, double *__restrict__ z, const CgScalars *sc)
{
    if (sc != nullptr && sc->done) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double di = fma(a, d[i], b * inv_diag[i] * (r[i] - t[i]));
        d[i] = di;
        z[i] += di;
    }
}
// *out += x . y over the owned range
__global__ void dot_kernel(uint32_t n, const double *__restrict__ x, const double *__restrict__ y, double *out, const CgScalars *sc)
{
    if (sc != nullptr && sc->done) return;
    double s = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s = fma(x[i], y[i], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) atomicAdd(out, s);
    }
}
// y = D^-1 x / scale  (power iteration), norms through dot_kernel
__global__ void scale_diag_kernel(uint32_t n, const double *__restrict__ x, const double *__restrict__ inv_diag, const double *norm2, double *__restrict__ y)
{
    const double f = 1.0 / sqrt(*norm2);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = f * (inv_diag ? inv_diag[i] * x[i] : x[i]);
}

// ---- p-multigrid (polynomial global coarsening on the same cells; deal.II MGTransferGlobalCoarsening, the header the
//      reference already includes: CEED_bp/src/bp3.cc:27) ---------------------------------------------------------------
// One CTA per cell: the embedding of FE_Q(p_coarse) into FE_Q(p_fine) is the tensor product of the 1-D matrix
// P[jf][ic] = coarse Lagrange function ic at fine GLL node jf; continuous elements make every cell produce the same value at
// a shared fine DoF, so the cell contributions are weighted by 1 / valence (prolongate_and_add) and the restriction is the
// exact transpose (restrict_and_add).  Runtime sizes (nf, nc <= 9), shared-memory sweeps: a transfer moves ~2 vectors, the
// smoothers around it dozens.
__global__ void ptransfer_kernel(uint32_t n_cells, int nf, int nc, int transpose, const double *__restrict__ P, const double *__restrict__ wgt,
                                 const uint32_t *__restrict__ idx_f, const uint32_t *__restrict__ idx_c, double *fine, double *coarse,
                                 const CgScalars *sc)
{
    if (sc != nullptr && sc->done) return;
    extern __shared__ double psm[];
    const int nf3 = nf * nf * nf, nc3 = nc * nc * nc;
    double *sP = psm, *a0 = sP + nf * nc, *a1 = a0 + nf3;  // two ping-pong arrays of nf^3
    for (int t = threadIdx.x; t < nf * nc; t += blockDim.x) sP[t] = P[t];
    for (uint32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        __syncthreads();
        if (!transpose) {
            // coarse values [kc][jc][ic] -> a0
            for (int t = threadIdx.x; t < nc3; t += blockDim.x) {
                const uint32_t id = idx_c[(size_t)cell * nc3 + t];
                a0[t] = id == kInvalidIndex ? 0.0 : coarse[id];
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nc * nc * nf; t += blockDim.x) {  // x: a1[kc][jc][if]
                const int i = t % nf, r = t / nf;
                double v = 0.0;
                for (int ic = 0; ic < nc; ++ic) v = fma(sP[i * nc + ic], a0[r * nc + ic], v);
                a1[t] = v;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nc * nf * nf; t += blockDim.x) {  // y: a0[kc][jf][if]
                const int i = t % nf, j = (t / nf) % nf, k = t / (nf * nf);
                double v = 0.0;
                for (int jc = 0; jc < nc; ++jc) v = fma(sP[j * nc + jc], a1[(k * nc + jc) * nf + i], v);
                a0[t] = v;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nf3; t += blockDim.x) {  // z, weight, add
                const uint32_t id = idx_f[(size_t)cell * nf3 + t];
                if (id == kInvalidIndex) continue;
                const int ij = t % (nf * nf), k = t / (nf * nf);
                double v = 0.0;
                for (int kc = 0; kc < nc; ++kc) v = fma(sP[k * nc + kc], a0[kc * nf * nf + ij], v);
                atomicAdd(fine + id, wgt[id] * v);
            }
        } else {
            for (int t = threadIdx.x; t < nf3; t += blockDim.x) {  // weighted fine values -> a0[kf][jf][if]
                const uint32_t id = idx_f[(size_t)cell * nf3 + t];
                a0[t] = id == kInvalidIndex ? 0.0 : wgt[id] * fine[id];
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nc * nf * nf; t += blockDim.x) {  // z^T: a1[kc][jf][if]
                const int ij = t % (nf * nf), kc = t / (nf * nf);
                double v = 0.0;
                for (int k = 0; k < nf; ++k) v = fma(sP[k * nc + kc], a0[k * nf * nf + ij], v);
                a1[t] = v;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nc * nc * nf; t += blockDim.x) {  // y^T: a0[kc][jc][if]
                const int i = t % nf, jc = (t / nf) % nc, kc = t / (nf * nc);
                double v = 0.0;
                for (int j = 0; j < nf; ++j) v = fma(sP[j * nc + jc], a1[(kc * nf + j) * nf + i], v);
                a0[t] = v;
            }
            __syncthreads();
            for (int t = threadIdx.x; t < nc3; t += blockDim.x) {  // x^T, add
                const uint32_t id = idx_c[(size_t)cell * nc3 + t];
                if (id == kInvalidIndex) continue;
                const int ic = t % nc, r = t / nc;
                double v = 0.0;
                for (int i = 0; i < nf; ++i) v = fma(sP[i * nc + ic], a0[r * nf + i], v);
                atomicAdd(coarse + id, v);
            }
        }
    }
}
__global__ void valence_kernel(size_t n, const uint32_t *__restrict__ idx, double *cnt)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (idx[i] != kInvalidIndex) atomicAdd(cnt + idx[i], 1.0);
}
__global__ void reciprocal_kernel(uint32_t n, double *v)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = v[i] > 0.0 ? 1.0 / v[i] : 0.0;
}
// r = b - t ; x += e
__global__ void residual_kernel(uint32_t n, const double *__restrict__ b, const double *__restrict__ t, double *__restrict__ r, const CgScalars *sc)
{
    if (sc != nullptr && sc->done) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) r[i] = b[i] - t[i];
}
__global__ void add_kernel(uint32_t n, const double *__restrict__ e, double *__restrict__ x, const CgScalars *sc)
{
    if (sc != nullptr && sc->done) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] += e[i];
}

}  // namespace

struct CgWork {
    size_t n_local = 0;  // total length: components * (n_owned + n_ghost)
    double *r = nullptr, *p = nullptr, *v = nullptr, *xb = nullptr;  // xb: x and b for the host-buffer entry point
    double *cheb = nullptr;  // z | d | t of the Chebyshev preconditioner (3 vectors, allocated on first use)
    CgScalars *sc = nullptr;
    CgScalars *h_sc = nullptr;  // pinned
    ~CgWork()
    {
        cudaFree(r); cudaFree(p); cudaFree(v); cudaFree(xb); cudaFree(sc); cudaFree(cheb);
        if (h_sc) cudaFreeHost(h_sc);
    }
};

static int ensure_work(std::unique_ptr<CgWork> &w, size_t n_local, bool need_xb)
{
    if (!w || w->n_local != n_local) {
        w.reset();  // free the old workspace first (large problems: old + new may not fit together)
        // built locally and cached only when complete: a failed allocation (OOM) must not leave a workspace with null
        // vectors behind for the next call of the same size
        auto fresh = std::make_unique<CgWork>();
        fresh->n_local = n_local;
        const size_t bytes = sizeof(double) * std::max<size_t>(n_local, 1);
        B200FE_CUDA_TRY(cudaMalloc(&fresh->r, bytes));
        B200FE_CUDA_TRY(cudaMalloc(&fresh->p, bytes));
        B200FE_CUDA_TRY(cudaMalloc(&fresh->v, bytes));
        B200FE_CUDA_TRY(cudaMalloc(&fresh->sc, sizeof(CgScalars)));
        B200FE_CUDA_TRY(cudaMallocHost(&fresh->h_sc, sizeof(CgScalars)));
        w = std::move(fresh);
    }
    if (need_xb && !w->xb) B200FE_CUDA_TRY(cudaMalloc(&w->xb, 2 * sizeof(double) * std::max<size_t>(n_local, 1)));
    return B200FE_OK;
}

// one workspace per operator, keyed by the operator pointer (small table; operators are few).  The table itself is
// guarded; a solve on one operator from two host threads at once is not supported (the reference's operator is not
// re-entrant either).
static std::unique_ptr<CgWork> &work_of(Operator *op)
{
    static std::mutex mu;
    static std::list<std::pair<Operator *, std::unique_ptr<CgWork>>> table;  // list: references stay valid on growth
    std::lock_guard<std::mutex> lock(mu);
    for (auto &e : table)
        if (e.first == op) return e.second;
    table.emplace_back(op, nullptr);
    return table.back().second;
}

void cg_release_work(Operator *op)
{
    auto &w = work_of(op);
    w.reset();
}

struct ChebSpec {  // degree >= 1 terms of the polynomial in D^-1 A on [lambda_max / smoothing_range, lambda_max]
    int degree = 0;
    double lambda_max = 0.0, smoothing_range = 0.0;
};

// z = p_k(D^-1 A) D^-1 r  (three-term recurrence, Saad Alg. 12.1; k - 1 operator applications)
static int cheb_apply(Operator &op, double *z, double *d, double *t, const CgScalars *sc, const ChebSpec &c, const double *d_inv_diag,
                      const double *d_r, dim3 blocks, cudaStream_t s)
{
    const uint32_t n = op.n_owned;
    const double lmax = c.lambda_max, lmin = c.lambda_max / c.smoothing_range;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma1 = theta / delta;
    double rho = 1.0 / sigma1;
    cheb_first_kernel<<<blocks, 256, 0, s>>>(n, 1.0 / theta, d_r, d_inv_diag, d, z, sc);
    B200FE_CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    for (int k = 1; k < c.degree; ++k) {
        if (int rc = op_vmult(op, t, z, nullptr, true, true, s, 1)) return rc;
        const double rho_new = 1.0 / (2.0 * sigma1 - rho);
        cheb_step_kernel<<<blocks, 256, 0, s>>>(n, rho_new * rho, 2.0 * rho_new / delta, d_r, t, d_inv_diag, d, z, sc);
        B200FE_CUDA_TRY(cudaGetLastError());
        ++g_launch_count;
        rho = rho_new;
    }
    return B200FE_OK;
}

// general preconditioner of the CG loop: z = M^-1 r into w.cheb (first n_local doubles); `sc` lets its kernels skip once done
using PrecondFn = std::function<int(const double *d_r, double *d_z, const CgScalars *sc, cudaStream_t s)>;

static int cg_run(Operator &op, int ncomp, CgWork &w, double *d_x, const double *d_b, const double *d_inv_diag, double abs_tol,
                  double rel_tol, int max_it, int check_every, b200fe_cg_result *res, cudaStream_t s, const PrecondFn *cheb = nullptr)
{
    NvtxRange range("cg_solver");
    const uint32_t n = op.n_owned;
    const size_t stride = op.n_local();
    const unsigned bx = n == 0 ? 1u : std::min<unsigned>((n + 1023) / 1024, std::max(148u * 8u / (unsigned)ncomp, 148u));
    const dim3 blocks(bx, (unsigned)ncomp);
    const dim3 blocks_local(stride == 0 ? 1u : std::min<unsigned>((unsigned)((stride + 1023) / 1024), std::max(148u * 8u / (unsigned)ncomp, 148u)), (unsigned)ncomp);
    const int jacobi = d_inv_diag != nullptr || cheb != nullptr;  // (general preconditioner: rho = r.z as well, z = w.cheb)
    // the vector kernels fuse z = D^-1 r for Jacobi; with the polynomial preconditioner z is a vector of its own
    const double *fused_diag = cheb ? nullptr : d_inv_diag;
    if (check_every < 1) check_every = 1;
    B200FE_CUDA_TRY(cudaMemsetAsync(w.sc, 0, sizeof(CgScalars), s));
    // p and v carry ghost entries: start from a clean ghost segment
    B200FE_CUDA_TRY(cudaMemsetAsync(w.p, 0, sizeof(double) * stride * ncomp, s));
    cg_init_kernel<<<blocks, 256, 0, s>>>(n, stride, d_b, d_x, w.r, fused_diag, w.sc);
    B200FE_CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    auto precondition = [&]() -> int {  // Chebyshev only: z = M^-1 r, acc[2] = r.z
        if (int rc = (*cheb)(w.r, w.cheb, w.sc, s)) return rc;
        dot_kernel<<<blocks, 256, 0, s>>>(n, w.r, w.cheb, &w.sc->acc[2], w.sc);
        B200FE_CUDA_TRY(cudaGetLastError());
        ++g_launch_count;
        return B200FE_OK;
    };
    if (cheb)
        if (int rc = precondition()) return rc;
    if (op.halo)
        if (int rc = halo_allreduce_sum(*op.halo, w.sc->acc + 1, 2, s)) return rc;
    // Sequence: S [U V X S] [U V X S] ...  (S = ReductionControl check + rho bookkeeping, U = p update,
    // V = operator apply with fused p.Ap, X = x/r update with fused r.r).  Once S has set `done`, U, V, X and S
    // are no-ops (V through KArgs::skip), so a chunk of iterations can be replayed blindly.
    auto poll = [&]() -> int {
        B200FE_CUDA_TRY(cudaMemcpyAsync(w.h_sc, w.sc, sizeof(CgScalars), cudaMemcpyDeviceToHost, s));
        B200FE_CUDA_TRY(cudaStreamSynchronize(s));
        if (op.halo && op.halo->p2p)  // a peer that never answered (bounded waits of the one-sided transport): fail, do not spin
            if (int rc = p2p_status(*op.halo)) return rc;
        return B200FE_OK;
    };
    auto scalar_step = [&]() -> int {
        cg_scalar_kernel<<<1, 1, 0, s>>>(w.sc, jacobi, max_it, abs_tol, rel_tol);
        B200FE_CUDA_TRY(cudaGetLastError());
        ++g_launch_count;
        return B200FE_OK;
    };
    auto iteration = [&]() -> int {
        cg_update_p_kernel<<<blocks_local, 256, 0, s>>>(n, (uint32_t)stride, stride, cheb ? w.cheb : w.r, fused_diag, w.p, d_x, w.v, w.sc, op.d_excl_mask);
        B200FE_CUDA_TRY(cudaGetLastError());
        ++g_launch_count;
        // block-diagonal operator: the fused p.Ap of every component lands in acc[0]; v was cleared by the p pass
        if (int rc = op_vmult(op, w.v, w.p, &w.sc->acc[0], true, true, s, ncomp, true)) return rc;
        if (op.halo)
            if (int rc = halo_allreduce_sum(*op.halo, w.sc->acc, 1, s)) return rc;
        cg_update_xr_kernel<<<blocks, 256, 0, s>>>(n, stride, w.v, fused_diag, w.r, w.sc);
        B200FE_CUDA_TRY(cudaGetLastError());
        ++g_launch_count;
        if (cheb)
            if (int rc = precondition()) return rc;
        if (op.halo)
            if (int rc = halo_allreduce_sum(*op.halo, w.sc->acc + 1, 2, s)) return rc;
        return scalar_step();
    };
    struct SkipGuard {  // the apply kernels watch sc->done only for the duration of this solve
        Operator &op;
        SkipGuard(Operator &o, const int *flag) : op(o) { op.d_skip = flag; }
        ~SkipGuard() { op.d_skip = nullptr; }
    } skip_guard(op, &w.sc->done);

    if (int rc = scalar_step()) return rc;
    if (int rc = poll()) return rc;
    int launched = 0;
    // CUDA-graph replay of `chunk` iterations at a time (single GPU, no per-launch event timing): the latency
    // floor of small problems is launch overhead (SURVEY.md hard part 4 / appendix B); the first iteration runs
    // directly so that every kernel is configured before capture.
    static const bool graphs_on = [] { const char *e = std::getenv("B200FE_CG_GRAPH"); return !e || std::atoi(e) != 0; }();
    const int chunk = check_every < 32 ? check_every : 32;
    // (stream capture is not permitted on the legacy / per-thread default streams: callers on those keep plain launches)
    const bool capturable = s != nullptr && s != cudaStreamLegacy && s != cudaStreamPerThread;
    const bool use_graph = graphs_on && capturable && !op.halo && !op.timing && chunk >= 2 && max_it >= 2 * chunk;
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches_per_chunk = 0;
    int rc_loop = B200FE_OK;
    while (!w.h_sc->done) {
        if (use_graph && launched >= 1 && max_it - launched >= chunk) {
            if (exec == nullptr) {
                cudaGraph_t graph = nullptr;
                const unsigned long long before = g_launch_count;
                B200FE_CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
                for (int k = 0; k < chunk && rc_loop == B200FE_OK; ++k) rc_loop = iteration();
                cudaError_t ce = cudaStreamEndCapture(s, &graph);
                launches_per_chunk = g_launch_count - before;
                g_launch_count = before;  // capture enqueued nothing; replays are counted below
                if (rc_loop != B200FE_OK) { if (graph) cudaGraphDestroy(graph); break; }
                if (ce != cudaSuccess) { rc_loop = fail_cuda(ce, "cudaStreamEndCapture"); break; }
                ce = cudaGraphInstantiate(&exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) { rc_loop = fail_cuda(ce, "cudaGraphInstantiate"); break; }
            }
            cudaError_t ce = cudaGraphLaunch(exec, s);
            if (ce != cudaSuccess) { rc_loop = fail_cuda(ce, "cudaGraphLaunch"); break; }
            g_launch_count += launches_per_chunk;
            launched += chunk;
            if ((rc_loop = poll()) != B200FE_OK) break;
        } else {
            if ((rc_loop = iteration()) != B200FE_OK) break;
            ++launched;
            if (launched % check_every == 0 || launched >= max_it || (use_graph && launched == 1))
                if ((rc_loop = poll()) != B200FE_OK) break;
        }
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (rc_loop != B200FE_OK) return rc_loop;
    cg_flush_x_kernel<<<blocks, 256, 0, s>>>(n, stride, w.p, d_x, w.sc);
    cg_clear_pending_kernel<<<1, 1, 0, s>>>(w.sc);
    B200FE_CUDA_TRY(cudaGetLastError());
    g_launch_count += 2;
    if (res) {
        res->iterations = w.h_sc->its;
        res->converged = w.h_sc->converged;
        res->initial_residual = w.h_sc->res0;
        res->final_residual = w.h_sc->res;
    }
    return w.h_sc->converged ? B200FE_OK : B200FE_ERR_NO_CONVERGENCE;
}

// ---- p-multigrid objects ------------------------------------------------------------------------------------------------
struct PTransfer {
    int nf = 0, nc = 0;
    uint32_t n_cells = 0, n_fine = 0, n_coarse = 0;
    const uint32_t *d_idx_f = nullptr, *d_idx_c = nullptr;  // borrowed index tables of the two levels
    double *d_P = nullptr, *d_w = nullptr;                  // owned: 1-D embedding matrix, 1 / valence of the fine DoFs
    ~PTransfer() { cudaFree(d_P); cudaFree(d_w); }
};

static int ptransfer_apply(const PTransfer &t, bool transpose, double *fine, double *coarse, const CgScalars *sc, cudaStream_t s)
{
    if (t.n_cells == 0) return B200FE_OK;
    const size_t smem = sizeof(double) * ((size_t)t.nf * t.nc + 2 * (size_t)t.nf * t.nf * t.nf);
    ptransfer_kernel<<<std::min<uint32_t>(t.n_cells, 148u * 8u), 128, smem, s>>>(t.n_cells, t.nf, t.nc, transpose ? 1 : 0, t.d_P, t.d_w, t.d_idx_f,
                                                                              t.d_idx_c, fine, coarse, sc);
    B200FE_CUDA_TRY(cudaGetLastError());
    ++g_launch_count;
    return B200FE_OK;
}

struct PmgLevel {
    Operator *op = nullptr;
    const double *inv_diag = nullptr;  // borrowed
    double lambda_max = 0.0;
    double *buf = nullptr;  // x | b | r | z | d | t, n_local each (owned)
    ~PmgLevel() { cudaFree(buf); }
};
struct Pmg {
    std::vector<std::unique_ptr<PmgLevel>> level;  // 0 = finest
    std::vector<const PTransfer *> transfer;       // transfer[l]: level l (fine) <-> level l + 1 (coarse); borrowed
    int degree = 3, coarse_degree = 8;
    double range = 20.0;
};

// x_l = V(b_l), x_l = 0 on entry: Chebyshev pre-smoothing, coarse correction, Chebyshev post-smoothing (symmetric)
static int pmg_vcycle_level(Pmg &m, size_t l, const CgScalars *sc, cudaStream_t s)
{
    PmgLevel &L = *m.level[l];
    Operator &op = *L.op;
    const uint32_t n = op.n_owned;
    const size_t nl = op.n_local();
    double *x = L.buf, *b = x + nl, *r = b + nl, *z = r + nl, *d = z + nl, *t = d + nl;
    const dim3 blocks(n == 0 ? 1u : std::min<unsigned>((n + 1023) / 1024, 148u * 8u));
    const bool coarsest = l + 1 == m.level.size();
    const ChebSpec spec{coarsest ? m.coarse_degree : m.degree, L.lambda_max, m.range};
    // pre-smoothing from x = 0: x = p(D^-1 A) D^-1 b
    if (int rc = cheb_apply(op, x, d, t, sc, spec, L.inv_diag, b, blocks, s)) return rc;
    if (coarsest) return B200FE_OK;
    PmgLevel &C = *m.level[l + 1];
    const size_t nlc = C.op->n_local();
    double *xc = C.buf, *bc = xc + nlc;
    // r = b - A x ; b_coarse = P^T r
    if (int rc = op_vmult(op, t, x, nullptr, true, true, s, 1)) return rc;
    residual_kernel<<<blocks, 256, 0, s>>>(n, b, t, r, sc);
    B200FE_CUDA_TRY(cudaMemsetAsync(bc, 0, sizeof(double) * nlc, s));
    if (int rc = ptransfer_apply(*m.transfer[l], true, r, bc, sc, s)) return rc;
    if (int rc = pmg_vcycle_level(m, l + 1, sc, s)) return rc;
    // x += P x_coarse
    if (int rc = ptransfer_apply(*m.transfer[l], false, x, xc, sc, s)) return rc;
    // post-smoothing: x += p(D^-1 A) D^-1 (b - A x)
    if (int rc = op_vmult(op, t, x, nullptr, true, true, s, 1)) return rc;
    residual_kernel<<<blocks, 256, 0, s>>>(n, b, t, r, sc);
    if (int rc = cheb_apply(op, z, d, t, sc, spec, L.inv_diag, r, blocks, s)) return rc;
    add_kernel<<<blocks, 256, 0, s>>>(n, z, x, sc);
    B200FE_CUDA_TRY(cudaGetLastError());
    g_launch_count += 3;
    return B200FE_OK;
}

static int pmg_apply(Pmg &m, const double *d_r, double *d_z, const CgScalars *sc, cudaStream_t s)
{
    PmgLevel &L = *m.level[0];
    const size_t nl = L.op->n_local();
    B200FE_CUDA_TRY(cudaMemcpyAsync(L.buf + nl, d_r, sizeof(double) * L.op->n_owned, cudaMemcpyDeviceToDevice, s));
    if (int rc = pmg_vcycle_level(m, 0, sc, s)) return rc;
    B200FE_CUDA_TRY(cudaMemcpyAsync(d_z, L.buf, sizeof(double) * L.op->n_owned, cudaMemcpyDeviceToDevice, s));
    return B200FE_OK;
}

}  // namespace b200fe

using namespace b200fe;

extern "C" {

int b200fe_cg_solve_components(b200fe_op *o, int n_components, double *d_x, const double *d_b, const double *d_inv_diag,
                               double abs_tol, double rel_tol, int max_it, int check_every, b200fe_cg_result *result, void *stream)
{
    B200FE_REQUIRE(o && d_x && d_b, "b200fe_cg_solve: null pointer");
    B200FE_REQUIRE(max_it >= 0, "b200fe_cg_solve: max_it < 0");
    B200FE_REQUIRE(n_components >= 1 && n_components <= 16, "b200fe_cg_solve: n_components outside 1..16");
    Operator &op = *reinterpret_cast<Operator *>(o);
    auto &w = work_of(&op);
    if (int rc = ensure_work(w, (size_t)op.n_local() * n_components, false)) return rc;
    int rc = cg_run(op, n_components, *w, d_x, d_b, d_inv_diag, abs_tol, rel_tol, max_it, check_every, result, (cudaStream_t)stream);
    if (rc == B200FE_ERR_NO_CONVERGENCE) fail(rc, "CG did not converge in %d iterations", max_it);
    return rc;
}

static int ensure_cheb(CgWork &w, size_t n_local, cudaStream_t s)
{
    if (!w.cheb) B200FE_CUDA_TRY(cudaMalloc(&w.cheb, 3 * sizeof(double) * std::max<size_t>(n_local, 1)));
    B200FE_CUDA_TRY(cudaMemsetAsync(w.cheb, 0, 3 * sizeof(double) * std::max<size_t>(n_local, 1), s));  // clean ghost segments
    return B200FE_OK;
}

int b200fe_cg_solve_chebyshev(b200fe_op *o, double *d_x, const double *d_b, const double *d_inv_diag, int degree, double lambda_max,
                              double smoothing_range, double abs_tol, double rel_tol, int max_it, int check_every,
                              b200fe_cg_result *result, void *stream)
{
    B200FE_REQUIRE(o && d_x && d_b && d_inv_diag, "b200fe_cg_solve_chebyshev: null pointer (the inverse diagonal is required)");
    B200FE_REQUIRE(degree >= 1 && degree <= 64, "b200fe_cg_solve_chebyshev: degree outside 1..64");
    B200FE_REQUIRE(lambda_max > 0.0 && smoothing_range > 1.0, "b200fe_cg_solve_chebyshev: need lambda_max > 0 and smoothing_range > 1");
    B200FE_REQUIRE(max_it >= 0, "b200fe_cg_solve_chebyshev: max_it < 0");
    Operator &op = *reinterpret_cast<Operator *>(o);
    auto &w = work_of(&op);
    if (int rc = ensure_work(w, op.n_local(), false)) return rc;
    if (int rc = ensure_cheb(*w, op.n_local(), (cudaStream_t)stream)) return rc;
    const ChebSpec spec{degree, lambda_max, smoothing_range};
    const uint32_t n = op.n_owned;
    const size_t nl = op.n_local();
    const dim3 blocks(n == 0 ? 1u : std::min<unsigned>((n + 1023) / 1024, 148u * 8u));
    CgWork *wk = w.get();
    const PrecondFn pre = [&, wk](const double *r, double *z, const CgScalars *sc, cudaStream_t st) -> int {
        return cheb_apply(op, z, wk->cheb + nl, wk->cheb + 2 * nl, sc, spec, d_inv_diag, r, blocks, st);
    };
    int rc = cg_run(op, 1, *w, d_x, d_b, d_inv_diag, abs_tol, rel_tol, max_it, check_every, result, (cudaStream_t)stream, &pre);
    if (rc == B200FE_ERR_NO_CONVERGENCE) fail(rc, "CG did not converge in %d iterations", max_it);
    return rc;
}

int b200fe_ptransfer_create(int p_fine, int p_coarse, uint32_t n_cells, const uint32_t *d_idx_fine, const uint32_t *d_idx_coarse,
                            uint32_t n_local_fine, uint32_t n_local_coarse, b200fe_ptransfer **out)
{
    B200FE_REQUIRE(out && (n_cells == 0 || (d_idx_fine && d_idx_coarse)), "b200fe_ptransfer_create: null pointer");
    B200FE_REQUIRE(p_coarse >= 1 && p_coarse < p_fine && p_fine <= 8, "b200fe_ptransfer_create: need 1 <= p_coarse < p_fine <= 8");
    auto t = std::make_unique<PTransfer>();
    t->nf = p_fine + 1; t->nc = p_coarse + 1; t->n_cells = n_cells; t->n_fine = n_local_fine; t->n_coarse = n_local_coarse;
    t->d_idx_f = d_idx_fine; t->d_idx_c = d_idx_coarse;
    // P[jf][ic] = coarse shape function ic at fine node jf = shape_values[ic*nf + jf] of FE_Q(p_coarse) on the GLL points of p_fine
    std::vector<double> sv((size_t)t->nc * t->nf), P((size_t)t->nf * t->nc);
    if (int rc = b200fe_basis_1d(p_coarse, t->nf, B200FE_QUAD_GLL, sv.data(), nullptr, nullptr, nullptr, nullptr)) return rc;
    for (int j = 0; j < t->nf; ++j)
        for (int i = 0; i < t->nc; ++i) P[(size_t)j * t->nc + i] = sv[(size_t)i * t->nf + j];
    B200FE_CUDA_TRY(cudaMalloc(&t->d_P, P.size() * sizeof(double)));
    B200FE_CUDA_TRY(cudaMemcpy(t->d_P, P.data(), P.size() * sizeof(double), cudaMemcpyHostToDevice));
    B200FE_CUDA_TRY(cudaMalloc(&t->d_w, std::max<size_t>(n_local_fine, 1) * sizeof(double)));
    B200FE_CUDA_TRY(cudaMemset(t->d_w, 0, std::max<size_t>(n_local_fine, 1) * sizeof(double)));
    const size_t n_idx = (size_t)n_cells * t->nf * t->nf * t->nf;
    if (n_idx) {
        valence_kernel<<<(unsigned)std::min<size_t>((n_idx + 255) / 256, 148u * 16u), 256>>>(n_idx, d_idx_fine, t->d_w);
        reciprocal_kernel<<<std::min<unsigned>((n_local_fine + 255) / 256, 148u * 16u), 256>>>(n_local_fine, t->d_w);
        B200FE_CUDA_TRY(cudaGetLastError());
        B200FE_CUDA_TRY(cudaDeviceSynchronize());
    }
    *out = reinterpret_cast<b200fe_ptransfer *>(t.release());
    return B200FE_OK;
}

void b200fe_ptransfer_destroy(b200fe_ptransfer *t) { delete reinterpret_cast<PTransfer *>(t); }

int b200fe_ptransfer_prolongate_add(b200fe_ptransfer *t, double *d_fine, const double *d_coarse, void *stream)
{
    B200FE_REQUIRE(t && d_fine && d_coarse, "b200fe_ptransfer_prolongate_add: null pointer");
    return ptransfer_apply(*reinterpret_cast<PTransfer *>(t), false, d_fine, const_cast<double *>(d_coarse), nullptr, (cudaStream_t)stream);
}

int b200fe_ptransfer_restrict_add(b200fe_ptransfer *t, double *d_coarse, const double *d_fine, void *stream)
{
    B200FE_REQUIRE(t && d_fine && d_coarse, "b200fe_ptransfer_restrict_add: null pointer");
    return ptransfer_apply(*reinterpret_cast<PTransfer *>(t), true, const_cast<double *>(d_fine), d_coarse, nullptr, (cudaStream_t)stream);
}

int b200fe_pmg_create(int n_levels, b200fe_op *const *ops, const double *const *d_inv_diag, const double *lambda_max,
                      b200fe_ptransfer *const *transfers, int smoother_degree, double smoothing_range, int coarse_degree, b200fe_pmg **out)
{
    B200FE_REQUIRE(out && ops && d_inv_diag && lambda_max && (n_levels == 1 || transfers), "b200fe_pmg_create: null pointer");
    B200FE_REQUIRE(n_levels >= 1 && n_levels <= 8, "b200fe_pmg_create: n_levels outside 1..8");
    B200FE_REQUIRE(smoother_degree >= 1 && coarse_degree >= 1 && smoothing_range > 1.0, "b200fe_pmg_create: bad smoother parameters");
    auto m = std::make_unique<Pmg>();
    m->degree = smoother_degree; m->coarse_degree = coarse_degree; m->range = smoothing_range;
    for (int l = 0; l < n_levels; ++l) {
        B200FE_REQUIRE(ops[l] && d_inv_diag[l] && lambda_max[l] > 0.0, "b200fe_pmg_create: level %d incomplete", l);
        auto L = std::make_unique<PmgLevel>();
        L->op = reinterpret_cast<Operator *>(ops[l]);
        B200FE_REQUIRE(L->op->halo == nullptr, "b200fe_pmg_create: single-rank operators only (the transfer weights are local valences)");
        L->inv_diag = d_inv_diag[l]; L->lambda_max = lambda_max[l];
        const size_t bytes = 6 * sizeof(double) * std::max<size_t>(L->op->n_local(), 1);
        B200FE_CUDA_TRY(cudaMalloc(&L->buf, bytes));
        B200FE_CUDA_TRY(cudaMemset(L->buf, 0, bytes));
        m->level.push_back(std::move(L));
        if (l + 1 < n_levels) {
            B200FE_REQUIRE(transfers[l], "b200fe_pmg_create: transfer %d missing", l);
            m->transfer.push_back(reinterpret_cast<const PTransfer *>(transfers[l]));
        }
    }
    for (int l = 0; l + 1 < n_levels; ++l) {
        const PTransfer &t = *m->transfer[l];
        B200FE_REQUIRE(t.nf == m->level[l]->op->nm && t.nc == m->level[l + 1]->op->nm && t.n_fine == m->level[l]->op->n_local() &&
                           t.n_coarse == m->level[l + 1]->op->n_local() && t.n_cells == m->level[l]->op->n_cells,
                       "b200fe_pmg_create: transfer %d does not fit levels %d and %d", l, l, l + 1);
    }
    *out = reinterpret_cast<b200fe_pmg *>(m.release());
    return B200FE_OK;
}

void b200fe_pmg_destroy(b200fe_pmg *m) { delete reinterpret_cast<Pmg *>(m); }

int b200fe_pmg_vcycle(b200fe_pmg *pm, double *d_z, const double *d_r, void *stream)
{
    B200FE_REQUIRE(pm && d_z && d_r, "b200fe_pmg_vcycle: null pointer");
    return pmg_apply(*reinterpret_cast<Pmg *>(pm), d_r, d_z, nullptr, (cudaStream_t)stream);
}

int b200fe_cg_solve_pmg(b200fe_pmg *pm, double *d_x, const double *d_b, double abs_tol, double rel_tol, int max_it, int check_every,
                        b200fe_cg_result *result, void *stream)
{
    B200FE_REQUIRE(pm && d_x && d_b, "b200fe_cg_solve_pmg: null pointer");
    B200FE_REQUIRE(max_it >= 0, "b200fe_cg_solve_pmg: max_it < 0");
    Pmg &m = *reinterpret_cast<Pmg *>(pm);
    Operator &op = *m.level[0]->op;
    auto &w = work_of(&op);
    if (int rc = ensure_work(w, op.n_local(), false)) return rc;
    if (int rc = ensure_cheb(*w, op.n_local(), (cudaStream_t)stream)) return rc;
    const PrecondFn pre = [&m](const double *r, double *z, const CgScalars *sc, cudaStream_t st) -> int { return pmg_apply(m, r, z, sc, st); };
    int rc = cg_run(op, 1, *w, d_x, d_b, m.level[0]->inv_diag, abs_tol, rel_tol, max_it, check_every, result, (cudaStream_t)stream, &pre);
    if (rc == B200FE_ERR_NO_CONVERGENCE) fail(rc, "CG did not converge in %d iterations", max_it);
    return rc;
}

int b200fe_op_estimate_max_eigenvalue(b200fe_op *o, const double *d_inv_diag, int n_iterations, double *lambda_max, void *stream)
{
    B200FE_REQUIRE(o && lambda_max, "b200fe_op_estimate_max_eigenvalue: null pointer");
    B200FE_REQUIRE(n_iterations >= 1 && n_iterations <= 1000, "b200fe_op_estimate_max_eigenvalue: n_iterations outside 1..1000");
    Operator &op = *reinterpret_cast<Operator *>(o);
    cudaStream_t s = (cudaStream_t)stream;
    auto &w = work_of(&op);
    if (int rc = ensure_work(w, op.n_local(), false)) return rc;
    if (int rc = ensure_cheb(*w, op.n_local(), s)) return rc;
    const uint32_t n = op.n_owned;
    const size_t nl = op.n_local();
    double *y = w->cheb, *t = w->cheb + nl, *u = w->cheb + 2 * nl;  // iterate (unit norm), A y, D^-1 A y
    const unsigned bx = n == 0 ? 1u : std::min<unsigned>((n + 1023) / 1024, 148u * 8u);
    double *acc = &w->sc->acc[0];
    // start vector: deterministic, all frequencies (deal.II starts its eigenvalue CG from a non-constant vector as well)
    std::vector<double> h(n);
    uint64_t g = op.halo ? 1 + (uint64_t)op.halo->rank * 0x9E3779B97F4A7C15ull : 1;
    for (uint32_t i = 0; i < n; ++i) {
        g = g * 6364136223846793005ull + 1442695040888963407ull;
        h[i] = (double)(g >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
    B200FE_CUDA_TRY(cudaMemcpyAsync(u, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice, s));
    double lam = 0.0;
    for (int it = 0; it <= n_iterations; ++it) {
        // y = u / |u| ; t = A y ; u = D^-1 t ; lambda = y . u  (Rayleigh quotient of D^-1 A in the D-inner product up to scaling)
        B200FE_CUDA_TRY(cudaMemsetAsync(acc, 0, 3 * sizeof(double), s));
        dot_kernel<<<bx, 256, 0, s>>>(n, u, u, acc, nullptr);
        if (op.halo)
            if (int rc = halo_allreduce_sum(*op.halo, acc, 1, s)) return rc;
        scale_diag_kernel<<<bx, 256, 0, s>>>(n, u, nullptr, acc, y);
        B200FE_CUDA_TRY(cudaGetLastError());
        if (it == n_iterations) break;
        if (int rc = op_vmult(op, t, y, nullptr, true, true, s, 1)) return rc;
        B200FE_CUDA_TRY(cudaMemsetAsync(acc + 1, 0, sizeof(double), s));
        // u = D^-1 t (norm factor 1: acc[2] holds 1.0)
        const double one = 1.0;
        B200FE_CUDA_TRY(cudaMemcpyAsync(acc + 2, &one, sizeof(double), cudaMemcpyHostToDevice, s));
        scale_diag_kernel<<<bx, 256, 0, s>>>(n, t, d_inv_diag, acc + 2, u);
        dot_kernel<<<bx, 256, 0, s>>>(n, y, u, acc + 1, nullptr);
        B200FE_CUDA_TRY(cudaGetLastError());
        if (op.halo)
            if (int rc = halo_allreduce_sum(*op.halo, acc + 1, 1, s)) return rc;
        B200FE_CUDA_TRY(cudaMemcpyAsync(&lam, acc + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
        B200FE_CUDA_TRY(cudaStreamSynchronize(s));
    }
    *lambda_max = 1.2 * lam;  // deal.II's safety factor on the estimate (PreconditionChebyshev::estimate_eigenvalues)
    return B200FE_OK;
}

int b200fe_cg_solve(b200fe_op *o, double *d_x, const double *d_b, const double *d_inv_diag, double abs_tol,
                    double rel_tol, int max_it, int check_every, b200fe_cg_result *result, void *stream)
{
    return b200fe_cg_solve_components(o, 1, d_x, d_b, d_inv_diag, abs_tol, rel_tol, max_it, check_every, result, stream);
}

int b200fe_cg_solve_host(b200fe_op *o, double *h_x, const double *h_b, const double *d_inv_diag, double abs_tol,
                         double rel_tol, int max_it, int check_every, b200fe_cg_result *result, void *stream)
{
    B200FE_REQUIRE(o && h_x && h_b, "b200fe_cg_solve_host: null pointer");
    Operator &op = *reinterpret_cast<Operator *>(o);
    auto &w = work_of(&op);
    if (int rc = ensure_work(w, op.n_local(), true)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    double *d_x = w->xb, *d_b = w->xb + std::max<uint32_t>(op.n_local(), 1);
    // (the ghost entries of x and b are never read: CG works on the owned range, the operator on p and v)
    B200FE_CUDA_TRY(cudaMemcpyAsync(d_b, h_b, sizeof(double) * op.n_owned, cudaMemcpyHostToDevice, s));
    int rc = cg_run(op, 1, *w, d_x, d_b, d_inv_diag, abs_tol, rel_tol, max_it, check_every, result, s);
    if (rc != B200FE_OK && rc != B200FE_ERR_NO_CONVERGENCE) return rc;
    B200FE_CUDA_TRY(cudaMemcpyAsync(h_x, d_x, sizeof(double) * op.n_owned, cudaMemcpyDeviceToHost, s));
    B200FE_CUDA_TRY(cudaStreamSynchronize(s));
    if (rc == B200FE_ERR_NO_CONVERGENCE) fail(rc, "CG did not converge in %d iterations", max_it);
    return rc;
}

int b200fe_op_vmult_host(b200fe_op *o, double *h_dst, const double *h_src, void *stream)
{
    B200FE_REQUIRE(o && h_dst && h_src, "b200fe_op_vmult_host: null pointer");
    Operator &op = *reinterpret_cast<Operator *>(o);
    auto &w = work_of(&op);
    if (int rc = ensure_work(w, op.n_local(), false)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    B200FE_CUDA_TRY(cudaMemsetAsync(w->p, 0, sizeof(double) * op.n_local(), s));
    B200FE_CUDA_TRY(cudaMemcpyAsync(w->p, h_src, sizeof(double) * op.n_owned, cudaMemcpyHostToDevice, s));
    if (int rc = op_vmult(op, w->v, w->p, nullptr, true, true, s)) return rc;
    B200FE_CUDA_TRY(cudaMemcpyAsync(h_dst, w->v, sizeof(double) * op.n_owned, cudaMemcpyDeviceToHost, s));
    B200FE_CUDA_TRY(cudaStreamSynchronize(s));
    return B200FE_OK;
}

}  // extern "C"
