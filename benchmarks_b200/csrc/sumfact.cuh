// sumfact.cuh -- sum-factorised cell kernels for sm_100a (FP64).
//
// One templated kernel covers the reference's whole kernel zoo for this path:
//   E-vector BK1 / BK3 / BK5   CEED_BK/include/kernels/BK{1,3,5}/templated_cuda_kernels.cuh
//   L-vector BP apply          CEED_bp/include/bk3_kokkos_kernel.h:27-389 (nq = p+2)
//                              bakeoff_problems_dealii/include/bk3_kokkos_kernel.h:32-434 (nq = p+1)
//   Helmholtz quad op          bp5_kokkos/benchmark.cc:62-137
//
// Design (B200-first, not a port):
//   * a CTA works on EPB elements at once; thread <-> (element, q, r) of the quadrature plane,
//     the third (slowest, p) index lives in a register column;
//   * the 1-D matrices B and D travel as a __grid_constant__ kernel parameter, so every
//     contraction along the register column is a DFMA with a constant-bank operand
//     (no shared-memory or register cost for the matrix);
//   * in-plane contractions read shared memory with warp-broadcast column reads and
//     row reads; staging arrays that are read row-wise use odd row strides (bank-conflict free);
//   * geometric factors are streamed straight from HBM with coalesced loads, software-prefetched
//     one p-layer ahead; nothing else touches DRAM except the element vectors;
//   * the transposed derivative along p is accumulated in registers while the layer is hot
//     (w[n] += D[p][n] * rr), so only two of the three flux components go through shared memory;
//   * L-vector mode fuses masked gather, additive scatter (RED.ADD.F64) and the CG inner
//     product src.(A src) into the same kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifdef B200FE_EVEN_ODD
#include "eo_contract.h"
#endif

namespace b200fe {

constexpr uint32_t kInvalidIndex = 0xFFFFFFFFu;  // numbers::invalid_unsigned_int

enum : int { QOP_LAPLACE = 1, QOP_MASS = 2, QOP_HELMHOLTZ = 3,
             // geometry evaluated on the fly for affine cells: G(q) = cellG * w_p w_q w_r with six
             // per-cell constants instead of the streamed 6 nq^3 factors; sumfact2 kernel, L-vector operators only
             QOP_AFFINE = 4 };

// 1-D matrices in the BK layout: B[q*NM+i] (CEED_BK BK1 serial_kernels.hpp:39),
// D[p*NQ+n] = derivative of collocation function n at point p (BK3 serial_kernels.hpp:98).
template <int NM, int NQ>
struct Mats {
    double B[NQ * NM];
    double D[NQ * NQ];
    double W[NQ];  // 1-D quadrature weights (affine on-the-fly geometry only)
#ifdef B200FE_EVEN_ODD
    eo::EoMats<NM, NQ> E;  // even / odd halves of B and D (tuning variant, eo_contract.h)
#endif
};

struct KArgs {
    uint32_t n_elems;       // elements handled by this launch
    const double *G;        // [e][6][NQ^3], components rr,rs,rt,ss,st,tt, r <-> slowest index
    const double *JxW;      // [e][NQ^3]
    const double *in;       // E-vector [e][NM^3]  | L-vector src (owned + ghosts)
    double *out;            // E-vector [e][NM^3]  | L-vector dst
    const uint32_t *idx;    // L-vector only: [e][NM^3], kInvalidIndex = constrained
    double *dot;            // L-vector only, optional: += sum_e u_e . (A_e u_e)
    const double *cellG;    // affine geometry only: [e][8] = det J * K K^T (rr,rs,rt,ss,st,tt), det J, pad
    const int *skip;        // optional: when *skip != 0 the launch is a no-op (CG iterations replayed after convergence)
};

constexpr __host__ __device__ int odd(int n) { return n | 1; }
constexpr __host__ __device__ int cmax(int a, int b) { return a > b ? a : b; }

template <int NM, int NQ, bool COLL>
struct Layout {
    static constexpr int N2 = NQ * NQ, N3 = N2 * NQ, M3 = NM * NM * NM;
    static constexpr int RU = odd(NM);  // row stride of nodal staging arrays U, Z   [i][j][k]
    static constexpr int RA = odd(NQ);  // row stride of half-interpolated arrays A, Y [i][j][r]
    // one region must hold U (NM^2*RU), A (NM^2*RA), B~ (NM*NQ^2) or a quadrature array (NQ^3)
    static constexpr int REGION = COLL ? N3 : cmax(cmax(NM * NM * RU, NM * NM * RA), N3);
    static constexpr int PER_ELEM = 3 * REGION;
    static constexpr size_t smem_bytes(int epb) { return sizeof(double) * (size_t)(N2 + epb * PER_ELEM); }
};

__device__ __forceinline__ double ldg_stream(const double *p) { return __ldg(p); }

// ---------------------------------------------------------------------------------------------
template <int NM, int NQ, bool COLL, int QOP, bool LVEC, int EPB, int MINB>
__global__ void __launch_bounds__(EPB *NQ *NQ, MINB)
    sumfact_kernel(const __grid_constant__ Mats<NM, NQ> m, const KArgs a)
{
    using L = Layout<NM, NQ, COLL>;
    constexpr int N2 = L::N2, N3 = L::N3, M3 = L::M3, RU = L::RU, RA = L::RA;
    constexpr int T = EPB * N2;
    constexpr bool LAP = (QOP & QOP_LAPLACE) != 0, MASS = (QOP & QOP_MASS) != 0;
#ifdef B200FE_ROLL_P1
    constexpr bool ROLL_P1 = B200FE_ROLL_P1 != 0;
#else
    constexpr bool ROLL_P1 = NQ >= 7;  // rolled flux loop: ~120 registers instead of 220+ (2 CTAs/SM)
#endif
    static_assert(!COLL || NM == NQ, "collocated operators need nm == nq");

    extern __shared__ double smem[];
    double *sD = smem;  // copy of D for thread-dependent rows / columns
    const int tid = threadIdx.x;
    const int el = tid / N2;            // element slot inside the CTA
    const int t2 = tid - el * N2;       // q*NQ + r
    const int q = t2 / NQ, r = t2 - q * NQ;
    double *R0 = smem + N2 + el * L::PER_ELEM, *R1 = R0 + L::REGION, *R2 = R1 + L::REGION;

    if constexpr (LAP) {
        for (int i = tid; i < N2; i += T) sD[i] = m.D[i];
        __syncthreads();
    }

    double dot_acc = 0.0;
    const uint32_t n_batches = (a.n_elems + EPB - 1) / EPB;
    for (uint32_t eb = blockIdx.x; eb < n_batches; eb += gridDim.x) {
        const uint32_t e = eb * EPB + el;
        const bool active = e < a.n_elems;
        double v[NQ];  // register column over p at quadrature point (q, r)

        // -------------------------------------------------------------------------------
        // load + forward interpolation
        // -------------------------------------------------------------------------------
        constexpr int NK = (M3 + N2 - 1) / N2;   // nodal values per thread
        [[maybe_unused]] double u_keep[NK];      // gathered values (for the fused dot)
        [[maybe_unused]] uint32_t i_keep[NK];    // their indices (reused by the scatter)
        if constexpr (COLL) {
            // nodal values are the quadrature values: straight into the register column
            if constexpr (LVEC) {
#pragma unroll
                for (int p = 0; p < NQ; ++p) {
                    const uint32_t id = active ? __ldg(a.idx + (size_t)e * N3 + p * N2 + t2) : kInvalidIndex;
                    i_keep[p] = id;
                    v[p] = id == kInvalidIndex ? 0.0 : __ldg(a.in + id);
                    u_keep[p] = v[p];
                }
            } else {
#pragma unroll
                for (int p = 0; p < NQ; ++p) v[p] = active ? ldg_stream(a.in + (size_t)e * N3 + p * N2 + t2) : 0.0;
            }
        } else {
            // U -> R0  [i][j][k], row stride RU
            {
#pragma unroll
                for (int c = 0; c < NK; ++c) {
                    const int l = t2 + c * N2;
                    if (l >= M3) break;
                    double val;
                    if constexpr (LVEC) {
                        const uint32_t id = active ? __ldg(a.idx + (size_t)e * M3 + l) : kInvalidIndex;
                        i_keep[c] = id;
                        val = id == kInvalidIndex ? 0.0 : __ldg(a.in + id);
                        u_keep[c] = val;
                    } else {
                        val = active ? ldg_stream(a.in + (size_t)e * M3 + l) : 0.0;
                    }
                    R0[(l / NM) * RU + (l % NM)] = val;
                }
            }
            __syncthreads();
            // sweep k -> r : items (i,j); R0 rows -> R1  A[i][j][r], row stride RA
            if (t2 < NM * NM) {
                double u[NM];
#pragma unroll
                for (int k = 0; k < NM; ++k) u[k] = R0[t2 * RU + k];
#pragma unroll
                for (int rr = 0; rr < NQ; ++rr) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NM; ++k) s = fma(m.B[rr * NM + k], u[k], s);
                    R1[t2 * RA + rr] = s;
                }
            }
            __syncthreads();
            // sweep j -> q : items (i, r); R1 columns -> R2  B~[i][q][r] (dense)
            if (t2 < NM * NQ) {
                const int i = t2 / NQ, r2 = t2 - i * NQ;
                double u[NM];
#pragma unroll
                for (int j = 0; j < NM; ++j) u[j] = R1[(i * NM + j) * RA + r2];
#pragma unroll
                for (int qq = 0; qq < NQ; ++qq) {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < NM; ++j) s = fma(m.B[qq * NM + j], u[j], s);
                    R2[(i * NQ + qq) * NQ + r2] = s;
                }
            }
            __syncthreads();
            // sweep i -> p : own (q, r); R2 column -> registers
            {
                double u[NM];
#pragma unroll
                for (int i = 0; i < NM; ++i) u[i] = R2[i * N2 + t2];
#pragma unroll
                for (int p = 0; p < NQ; ++p) {
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < NM; ++i) s = fma(m.B[p * NM + i], u[i], s);
                    v[p] = s;
                }
            }
        }

        // -------------------------------------------------------------------------------
        // operator at the quadrature points: w = D^T G D v (+ JxW v)
        // -------------------------------------------------------------------------------
        double w[NQ];
#pragma unroll
        for (int p = 0; p < NQ; ++p) w[p] = 0.0;

        if constexpr (LAP) {
            const double *Ge = a.G + (size_t)(active ? e : 0) * 6 * N3 + t2;
            double g[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) g[c] = active ? ldg_stream(Ge + c * N3) : 0.0;

            // V -> R0 (dense [p][q][r])
#pragma unroll
            for (int p = 0; p < NQ; ++p) R0[p * N2 + t2] = v[p];
            double dq[NQ], dr[NQ];
#pragma unroll
            for (int n = 0; n < NQ; ++n) {
                dq[n] = sD[q * NQ + n];
                dr[n] = sD[r * NQ + n];
            }
            __syncthreads();
#pragma unroll(ROLL_P1 ? 1 : NQ)
            for (int p = 0; p < NQ; ++p) {
                double gn[6];
                if (p + 1 < NQ) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) gn[c] = active ? ldg_stream(Ge + c * N3 + (p + 1) * N2) : 0.0;
                }
                double qr = 0.0, qs = 0.0, qt = 0.0;
#pragma unroll
                for (int n = 0; n < NQ; ++n) {
                    qr = fma(m.D[p * NQ + n], v[n], qr);
                    qs = fma(dq[n], R0[p * N2 + n * NQ + r], qs);
                    qt = fma(dr[n], R0[p * N2 + q * NQ + n], qt);
                }
                const double fr = g[0] * qr + g[1] * qs + g[2] * qt;
                const double fs = g[1] * qr + g[3] * qs + g[4] * qt;
                const double ft = g[2] * qr + g[4] * qs + g[5] * qt;
#pragma unroll
                for (int n = 0; n < NQ; ++n) w[n] = fma(m.D[p * NQ + n], fr, w[n]);
                R1[p * N2 + t2] = fs;
                R2[p * N2 + t2] = ft;
                if (p + 1 < NQ) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) g[c] = gn[c];
                }
            }
#pragma unroll
            for (int n = 0; n < NQ; ++n) {  // columns of D for the transposed in-plane derivatives
                dq[n] = sD[n * NQ + q];
                dr[n] = sD[n * NQ + r];
            }
            __syncthreads();
#pragma unroll
            for (int p = 0; p < NQ; ++p) {
                double s = w[p];
#pragma unroll
                for (int n = 0; n < NQ; ++n) {
                    s = fma(dq[n], R1[p * N2 + n * NQ + r], s);
                    s = fma(dr[n], R2[p * N2 + q * NQ + n], s);
                }
                w[p] = s;
            }
        }
        if constexpr (MASS) {
            const double *Je = a.JxW + (size_t)(active ? e : 0) * N3 + t2;
#pragma unroll
            for (int p = 0; p < NQ; ++p) w[p] = fma(active ? ldg_stream(Je + p * N2) : 0.0, v[p], w[p]);
        }

        // -------------------------------------------------------------------------------
        // backward interpolation + store / scatter
        // -------------------------------------------------------------------------------
        if constexpr (COLL) {
            if constexpr (LVEC) {
#pragma unroll
                for (int p = 0; p < NQ; ++p) {
                    if (i_keep[p] != kInvalidIndex) {
                        atomicAdd(a.out + i_keep[p], w[p]);
                        dot_acc = fma(u_keep[p], w[p], dot_acc);
                    }
                }
            } else {
                if (active) {
#pragma unroll
                    for (int p = 0; p < NQ; ++p) a.out[(size_t)e * N3 + p * N2 + t2] = w[p];
                }
            }
        } else {
            // sweep p -> i in registers; X[i][q][r] -> R0 (dense).  R0 (V) was last read in the
            // flux loop, which every thread left through the barrier above.
#pragma unroll
            for (int i = 0; i < NM; ++i) {
                double s = 0.0;
#pragma unroll
                for (int p = 0; p < NQ; ++p) s = fma(m.B[p * NM + i], w[p], s);
                R0[i * N2 + t2] = s;
            }
            __syncthreads();
            // sweep q -> j : items (i, r); R0 columns -> R1  Y[i][j][r], row stride RA
            if (t2 < NM * NQ) {
                const int i = t2 / NQ, r2 = t2 - i * NQ;
                double x[NQ];
#pragma unroll
                for (int qq = 0; qq < NQ; ++qq) x[qq] = R0[(i * NQ + qq) * NQ + r2];
#pragma unroll
                for (int j = 0; j < NM; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int qq = 0; qq < NQ; ++qq) s = fma(m.B[qq * NM + j], x[qq], s);
                    R1[(i * NM + j) * RA + r2] = s;
                }
            }
            __syncthreads();
            // sweep r -> k : items (i, j); R1 rows -> R2  Z[i][j][k], row stride RU
            if (t2 < NM * NM) {
                double x[NQ];
#pragma unroll
                for (int rr = 0; rr < NQ; ++rr) x[rr] = R1[t2 * RA + rr];
#pragma unroll
                for (int k = 0; k < NM; ++k) {
                    double s = 0.0;
#pragma unroll
                    for (int rr = 0; rr < NQ; ++rr) s = fma(m.B[rr * NM + k], x[rr], s);
                    R2[t2 * RU + k] = s;
                }
            }
            __syncthreads();
            {
#pragma unroll
                for (int c = 0; c < NK; ++c) {
                    const int l = t2 + c * N2;
                    if (l >= M3) break;
                    const double z = R2[(l / NM) * RU + (l % NM)];
                    if constexpr (LVEC) {
                        if (i_keep[c] != kInvalidIndex) {
                            atomicAdd(a.out + i_keep[c], z);
                            dot_acc = fma(u_keep[c], z, dot_acc);
                        }
                    } else {
                        if (active) a.out[(size_t)e * M3 + l] = z;
                    }
                }
            }
            // next iteration's first smem write goes to R0; R2 readers are fenced by its first barrier
        }
    }

    if constexpr (LVEC) {
        if (a.dot != nullptr) {
            // block reduction of the fused inner product.  The CTA size is a multiple of the plane size,
            // not of 32, so the last warp is partial: no full-mask shuffles here (they would read lanes
            // that do not exist) -- shared-memory atomics instead, once per thread per launch.
            __shared__ double red;
            __syncthreads();
            if (tid == 0) red = 0.0;
            __syncthreads();
            atomicAdd(&red, dot_acc);
            __syncthreads();
            if (tid == 0) atomicAdd(a.dot, red);
        }
    }
}

}  // namespace b200fe
