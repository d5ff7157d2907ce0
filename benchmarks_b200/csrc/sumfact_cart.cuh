// sumfact_cart.cuh -- separable cell kernel for axis-aligned cells, any quadrature (L-vector Laplace operators).
//
// On a cell that is an axis-aligned box (deal.II MatrixFree's "cartesian" cell type; every mesh of the reference drivers:
// CEED_bp/src/bp3.cc:452-488 builds cubes) the geometric factors are diagonal and separable, G_aa(q) = c_aa w_p w_q w_r, so
// the cell matrix of  B^T D^T G D B  is a sum of three Kronecker products of two nm x nm matrices,
//
//      A_e = c_rr (K x M x M) + c_ss (M x K x M) + c_tt (M x M x K),     K = (D B)^T W (D B),   M = B^T W B,
//
// (r <-> slowest local index, as everywhere in this library).  The operator is then applied on the nm^3 nodal values directly:
// no interpolation to the nq^3 quadrature points, no geometric factors in memory, 7 one-dimensional contractions of length nm
// instead of 6 of mixed length nm -> nq plus 6 of length nq:
//
//      a = M_k u,  b = K_k u;    a2 = M_j a,  a3 = K_j a,  b2 = M_j b;    out = c_rr K_i a2 + M_i (c_ss a3 + c_tt b2).
//
// Same thread scheme as sumfact2.cuh: nm^2 threads per element acting in three layouts (register column over i, row over k,
// column over j), two staging arrays per element whose strides come from the same compile-time search; gathered inputs are
// software-pipelined one batch ahead; the scatter writes cell-interior DoFs with plain stores (KArgs::excl_interior); the CG
// inner product u.(A u) is accumulated from the nodal values.  The collocated case (M diagonal) has its own 3-contraction
// branch in sumfact2.cuh (QOP_CARTESIAN there); this kernel serves the interpolated operators (BP3: QGauss(p+2), "bp35").
//
// Numerics contract: <= 1e-12 relative max-norm against the oracle's B^T D^T G D B (tests/test_operator_gpu.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sumfact2.cuh"

namespace b200fe {

// Both 1-D matrices are symmetric AND point-symmetric (A(nm-1-i, nm-1-j) = A(i,j): real basis, symmetric points).  Every
// contraction therefore runs through the even-odd split (eo_contract.h), and the two half matrices are symmetric again, so
// they are stored packed (upper triangle): nm = 9 needs 15 + 10 numbers per matrix instead of 81.  Besides halving the
// multiply-adds this is what keeps the kernel in registers: ptxas keeps loop-invariant matrix entries that do not fit the
// uniform register file in ordinary registers -- with the full matrices the nm = 7, 9 builds wanted > 255 registers and
// spilled ~340 B/thread at the 168-register cap.
template <int NM>
struct CartMats {
    eo::SymEo<NM> K;  // 1-D stiffness matrix
    eo::SymEo<NM> M;  // 1-D mass matrix
};

namespace cart {
template <int NM>
struct LayoutC {
    static constexpr int N2 = NM * NM, M3 = N2 * NM;
    static constexpr v2::Strides SA = v2::best_strides(NM, NM);
    static constexpr int RA = SA.ra, PA = SA.pa;                // [i][j][k]: row stride (over j) and plane stride (over i)
    static constexpr int ARR = (NM * PA + 1) & ~1;              // one staging array (even: 16-byte alignment of the second)
    static constexpr int pad_elem(int w) { while ((w - N2) % 16 != 0) ++w; return w; }
    static constexpr int WORK_PER_ELEM = pad_elem(2 * ARR);     // element stride == nm^2 (mod 16), as in sumfact2
    static constexpr size_t smem_bytes(int epb) { return sizeof(double) * (size_t)epb * WORK_PER_ELEM; }
};

// out = A in  (packed even-odd halves in the constant bank, register column)
template <int NM>
__device__ __forceinline__ void mat_col(const eo::SymEo<NM> &A, const double (&in)[NM], double (&out)[NM])
{
    eo::sym_apply<NM>(A, in, out);
}
}  // namespace cart

// QOP: QOP_LAPLACE, QOP_MASS or QOP_HELMHOLTZ -- the mass term of an affine cell is separable as well, det J (M x M x M), so
// the Helmholtz operator costs no contraction more than the Laplacian (c_m a2 joins the sum under M_i) and the mass operator
// three in total.
template <int NM, int EPB, int MINB, int QOP = QOP_LAPLACE>
__global__ void __launch_bounds__(EPB *NM *NM, MINB) sumfact_cart_kernel(const __grid_constant__ CartMats<NM> m, const KArgs a)
{
    constexpr bool LAP = (QOP & QOP_LAPLACE) != 0, MASS = (QOP & QOP_MASS) != 0;
    using L = cart::LayoutC<NM>;
    constexpr int N2 = L::N2, M3 = L::M3, RA = L::RA, PA = L::PA;
    if (a.skip != nullptr && *a.skip != 0) return;  // uniform across the grid
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *work = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x;
    const int el = tid / N2, t2 = tid - el * N2;
    const int ta = t2 / NM, tb = t2 - ta * NM;
    double *X1 = work + el * L::WORK_PER_ELEM, *X2 = X1 + L::ARR;
    const uint32_t n_batches = (a.n_elems + EPB - 1) / EPB;

    double cur_val[NM], nxt_val[NM];
    uint32_t cur_idx[NM], nxt_idx[NM];
    auto load_idx = [&](uint32_t eb_, uint32_t (&ix)[NM]) {
        const uint32_t e_ = eb_ * EPB + el;
        const bool ok = eb_ < n_batches && e_ < a.n_elems;
#pragma unroll
        for (int n = 0; n < NM; ++n) ix[n] = ok ? __ldg(a.idx + (size_t)e_ * M3 + n * N2 + t2) : kInvalidIndex;  // plane n, position t2
    };
    auto load_val = [&](const uint32_t (&ix)[NM], double (&val)[NM]) {
#pragma unroll
        for (int n = 0; n < NM; ++n) val[n] = ix[n] == kInvalidIndex ? 0.0 : __ldg(a.in + ix[n]);
    };
    load_idx(blockIdx.x, cur_idx);
    load_val(cur_idx, cur_val);
    // layout P: this thread's two fixed local indices (j, k) are interior ones (KArgs::excl_interior)
    const bool excl_t2 = a.excl_interior != 0 && ta >= 1 && ta <= NM - 2 && tb >= 1 && tb <= NM - 2;

    double dot_acc = 0.0;
    for (uint32_t eb = blockIdx.x; eb < n_batches; eb += gridDim.x) {
        const uint32_t e = eb * EPB + el;
        const bool active = e < a.n_elems;
        const uint32_t nb = eb + gridDim.x;
        load_idx(nb, nxt_idx);  // next batch's index rows; the dependent gathers are issued after the first sweep
        const double *c8 = a.cellG + (size_t)(active ? e : 0) * 8;
        [[maybe_unused]] const double c_rr = (LAP && active) ? __ldg(c8 + 0) : 0.0, c_ss = (LAP && active) ? __ldg(c8 + 3) : 0.0,
                                      c_tt = (LAP && active) ? __ldg(c8 + 5) : 0.0;
        [[maybe_unused]] const double c_m = (MASS && active) ? __ldg(c8 + 6) : 0.0;  // det J

        // layout P: thread (j,k) = (ta,tb) owns the column over i -> X1[i][j][k]
#pragma unroll
        for (int i = 0; i < NM; ++i) X1[i * PA + ta * RA + tb] = cur_val[i];
        __syncthreads();
        {   // layout R: thread (i,j) = (ta,tb), row over k:  a = M u -> X1 (in place),  b = K u -> X2
            double u[NM], o[NM];
#pragma unroll
            for (int k = 0; k < NM; ++k) u[k] = X1[ta * PA + tb * RA + k];
            cart::mat_col<NM>(m.M, u, o);
#pragma unroll
            for (int k = 0; k < NM; ++k) X1[ta * PA + tb * RA + k] = o[k];
            if constexpr (LAP) {
                cart::mat_col<NM>(m.K, u, o);
#pragma unroll
                for (int k = 0; k < NM; ++k) X2[ta * PA + tb * RA + k] = o[k];
            }
        }
        __syncthreads();
        load_val(nxt_idx, nxt_val);  // next batch's gathers (indices arrived during the first sweep)
        {   // layout Q: thread (i,k) = (ta,tb), column over j:  a2 = M a -> X1,  t = c_ss K a + c_tt M b -> X2
            double x[NM], o[NM];
            if constexpr (LAP) {
                double t[NM];
#pragma unroll
                for (int j = 0; j < NM; ++j) x[j] = X2[ta * PA + j * RA + tb];
                cart::mat_col<NM>(m.M, x, o);
#pragma unroll
                for (int j = 0; j < NM; ++j) o[j] *= c_tt;
#pragma unroll
                for (int j = 0; j < NM; ++j) x[j] = X1[ta * PA + j * RA + tb];
                cart::mat_col<NM>(m.K, x, t);
#pragma unroll
                for (int j = 0; j < NM; ++j) t[j] = fma(c_ss, t[j], o[j]);
                cart::mat_col<NM>(m.M, x, o);
#pragma unroll
                for (int j = 0; j < NM; ++j) {
                    X2[ta * PA + j * RA + tb] = MASS ? fma(c_m, o[j], t[j]) : t[j];  // Helmholtz: + det J * M_j M_k u under M_i
                    X1[ta * PA + j * RA + tb] = o[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < NM; ++j) x[j] = X1[ta * PA + j * RA + tb];
                cart::mat_col<NM>(m.M, x, o);
#pragma unroll
                for (int j = 0; j < NM; ++j) X1[ta * PA + j * RA + tb] = o[j];
            }
        }
        __syncthreads();
        {   // layout P again: out = c_rr K a2 + M t along i; scatter
            double x[NM], o2[NM];
            [[maybe_unused]] double o1[NM];
#pragma unroll
            for (int i = 0; i < NM; ++i) x[i] = X1[i * PA + ta * RA + tb];
            if constexpr (LAP) {
                cart::mat_col<NM>(m.K, x, o1);
#pragma unroll
                for (int i = 0; i < NM; ++i) x[i] = X2[i * PA + ta * RA + tb];
            }
            cart::mat_col<NM>(m.M, x, o2);
#pragma unroll
            for (int i = 0; i < NM; ++i) {
                const double w = LAP ? fma(c_rr, o1[i], o2[i]) : c_m * o2[i];
                dot_acc = fma(cur_val[i], w, dot_acc);  // u.(A u) from the nodal values (masked entries carry u = 0)
                if (cur_idx[i] != kInvalidIndex) {
                    if (excl_t2 && i >= 1 && i <= NM - 2) a.out[cur_idx[i]] = w;  // cell-interior DoF: sole writer
                    else atomicAdd(a.out + cur_idx[i], w);
                }
            }
        }
        // (the next batch stores into the slots of X1 this thread has just read: its own, no barrier needed; X2 is written
        //  again only after the barrier that follows those stores)
#pragma unroll
        for (int n = 0; n < NM; ++n) {
            cur_val[n] = nxt_val[n];
            cur_idx[n] = nxt_idx[n];
        }
    }

    if (a.dot != nullptr) {  // block reduction of the fused inner product (partial last warp: shared-memory atomics)
        __shared__ double red;
        __syncthreads();
        if (tid == 0) red = 0.0;
        __syncthreads();
        atomicAdd(&red, dot_acc);
        __syncthreads();
        if (tid == 0) atomicAdd(a.dot, red);
    }
}

}  // namespace b200fe
