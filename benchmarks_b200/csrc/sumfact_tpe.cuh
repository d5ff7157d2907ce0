// sumfact_tpe.cuh -- thread-per-element cell kernel for the tiny elements (nq <= 3: p = 1 everywhere, p = 2 collocated).
//
// Why (ncu, profiles/r02j_p1_kernels_summary.txt): with 4 or 9 threads per element the plane-per-thread kernel of
// sumfact2.cuh spends its time in shared memory -- 44-52 % of its wavefronts are bank conflicts (the geometric factors of
// consecutive elements sit 48 or 162 doubles apart, i.e. in the same or neighbouring banks), every 1-D sweep needs a CTA
// barrier for a handful of flops, and the half-warps of the partial sweeps are mostly idle: 0.60 of the HBM roofline at
// p = 1, the only BK3 / BP5 degrees below the north-star fraction.  An element of 8 or 27 nodal values fits the
// registers of ONE thread:
//   * thread = element: gather / load its nm^3 values, all three interpolation sweeps, the derivatives at its nq^3 points,
//     the flux, the transposed sweeps and the scatter happen in registers with compile-time indices -- no work arrays in
//     shared memory, no barrier inside an element;
//   * geometric factors: every thread issues ONE bulk copy (TMA, cp.async.bulk) of its own element's block
//     [6][nq^3] (384 or 1296 B) into its slot of the CTA's buffer, all completing on one mbarrier; slots are 2 (mod 16)
//     doubles apart, so the 16 lanes of a half-warp read 8 distinct banks (2-way instead of 16-way);
//   * the next batch's indices / values and the next batch's G copy are in flight behind the current element's arithmetic
//     (same software pipeline as sumfact2.cuh).
//   * JxW (mass, Helmholtz) is staged the same way: one contiguous bulk copy per batch when nq^3 is odd (slot stride 27:
//     conflict-free), per-thread copies into padded slots when it is even; a partial tail batch copies with plain loads.
// Same KArgs, same numerics contract (<= 1e-12 against the oracle) and the same fused epilogues (RED scatter, p.Ap) as the
// plane-per-thread kernel; stored geometric factors and / or JxW (Laplace, mass, Helmholtz).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sumfact2.cuh"

namespace b200fe {
namespace tpe {

// slot stride of one element's G block in shared memory: >= 6 nq^3, == 2 (mod 16) doubles (8 distinct banks per half-warp,
// 16-byte aligned for the bulk copy)
constexpr int g_slot(int nq)
{
    int s = 6 * nq * nq * nq;
    while (s % 16 != 2) ++s;
    return s;
}

constexpr int j_slot(int n3)
{
    int s = n3;
    while (s % 16 != 2) ++s;
    return s;
}

template <int NM, int NQ, bool COLL, int QOP, int TPB>
struct LayoutT {
    static constexpr int N3 = NQ * NQ * NQ, M3 = NM * NM * NM;
    static constexpr bool LAP = (QOP & QOP_LAPLACE) != 0;
    static constexpr bool MASS = (QOP & QOP_MASS) != 0;
    static constexpr int SLOT = LAP ? g_slot(NQ) : 0;
    // JxW: nq^3 odd -> the batch is one contiguous bulk copy and the odd slot stride is conflict-free as it is;
    //      nq^3 even -> per-thread copies into slots 2 (mod 16) apart, like G
    static constexpr bool J_BATCH_COPY = (N3 % 2) == 1;
    static constexpr int SLOT_J = !MASS ? 0 : (J_BATCH_COPY ? N3 : j_slot(N3));
    static constexpr size_t smem_bytes() { return 16 + sizeof(double) * (size_t)TPB * (SLOT + SLOT_J); }
};

}  // namespace tpe

template <int NM, int NQ, bool COLL, int QOP, bool LVEC, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) sumfact_tpe_kernel(const __grid_constant__ Mats<NM, NQ, false> m, const KArgs a)
{
    using L = tpe::LayoutT<NM, NQ, COLL, QOP, TPB>;
    constexpr int N3 = L::N3, M3 = L::M3, SLOT = L::SLOT, SLOT_J = L::SLOT_J;
    constexpr bool LAP = L::LAP, MASS = L::MASS, STAGED = LAP || MASS;
    static_assert(TPB % 2 == 0, "even CTA size: a full batch of JxW is a multiple of 16 bytes");
    static_assert(!COLL || NM == NQ, "collocated operators need nm == nq");
    static_assert(NQ <= 3, "thread-per-element kernel: nq <= 3");
    static_assert(!(QOP & QOP_AFFINE), "thread-per-element kernel: stored geometric factors");

    if (a.skip != nullptr && *a.skip != 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *Gs = reinterpret_cast<double *>(smem_raw + 16);
    const int tid = threadIdx.x;
    const double *Ge = Gs + tid * SLOT;
    double *Js = Gs + TPB * SLOT;
    const double *Je = Js + tid * SLOT_J;
    const uint32_t n_batches = (a.n_elems + TPB - 1) / TPB;

    auto issue_g = [&](uint32_t eb) {  // every thread fetches its own element's block(s); thread 0 announces the total
        const uint32_t first = eb * TPB;
        const uint32_t cnt = (a.n_elems - first) < (uint32_t)TPB ? (a.n_elems - first) : (uint32_t)TPB;
        constexpr uint32_t GB = 6 * N3 * sizeof(double), JB = N3 * sizeof(double);
        const bool j_tma = MASS && (!L::J_BATCH_COPY || cnt == (uint32_t)TPB);
        if (tid == 0) v2::mbar_expect_tx(bar, (LAP ? cnt * GB : 0u) + (j_tma ? cnt * JB : 0u));
        if constexpr (LAP)
            if ((uint32_t)tid < cnt) v2::bulk_g2s(Gs + tid * SLOT, a.G + (size_t)(first + tid) * 6 * N3, GB, bar);
        if constexpr (MASS) {
            if constexpr (L::J_BATCH_COPY) {
                if (j_tma) {
                    if (tid == 0) v2::bulk_g2s(Js, a.JxW + (size_t)first * N3, (uint32_t)TPB * JB, bar);
                } else if ((uint32_t)tid < cnt) {  // partial tail batch (its byte count need not be a multiple of 16): plain copy
                    for (int l = 0; l < N3; ++l) Js[tid * SLOT_J + l] = __ldg(a.JxW + (size_t)(first + tid) * N3 + l);
                }
            } else if ((uint32_t)tid < cnt)
                v2::bulk_g2s(Js + tid * SLOT_J, a.JxW + (size_t)(first + tid) * N3, JB, bar);
        }
    };
    if constexpr (STAGED) {
        if (tid == 0) {
            v2::mbar_init(bar, 1);
            v2::fence_mbar_init();
        }
        // unused slots of a tail batch: zeros, never NaNs (0 * stale in p.Ap)
        for (int i = tid; i < TPB * (SLOT + SLOT_J); i += TPB) Gs[i] = 0.0;
        v2::fence_proxy_async();
        __syncthreads();
        if (blockIdx.x < n_batches) issue_g(blockIdx.x);
    }
    uint32_t parity = 0;

    // inputs one batch ahead: the element's index row (L-vector) and, while they fit the register file, its nodal values
    constexpr bool PREFETCH_VAL = M3 <= 8;
    double cur[M3];
    [[maybe_unused]] double nxt[PREFETCH_VAL ? M3 : 1];
    [[maybe_unused]] uint32_t cur_idx[M3], nxt_idx[M3];
    auto load_idx = [&](uint32_t eb_, uint32_t (&ix)[M3]) {
        const uint32_t e_ = eb_ * TPB + tid;
        const bool ok = eb_ < n_batches && e_ < a.n_elems;
#pragma unroll
        for (int l = 0; l < M3; ++l) ix[l] = ok ? __ldg(a.idx + (size_t)e_ * M3 + l) : kInvalidIndex;
    };
    auto load_val = [&](uint32_t eb_, const uint32_t (&ix)[M3], double (&val)[M3]) {
        if constexpr (LVEC) {
#pragma unroll
            for (int l = 0; l < M3; ++l) val[l] = ix[l] == kInvalidIndex ? 0.0 : __ldg(a.in + ix[l]);
        } else {
            const uint32_t e_ = eb_ * TPB + tid;
            const bool ok = eb_ < n_batches && e_ < a.n_elems;
#pragma unroll
            for (int l = 0; l < M3; ++l) val[l] = ok ? __ldg(a.in + (size_t)e_ * M3 + l) : 0.0;
        }
    };
    if constexpr (LVEC) load_idx(blockIdx.x, cur_idx);
    load_val(blockIdx.x, cur_idx, cur);

    double dot_acc = 0.0;
    for (uint32_t eb = blockIdx.x; eb < n_batches; eb += gridDim.x) {
        const uint32_t e = eb * TPB + tid;
        const bool active = e < a.n_elems;
        const uint32_t nb = eb + gridDim.x;
        if constexpr (LVEC) load_idx(nb, nxt_idx);
        else if constexpr (PREFETCH_VAL) load_val(nb, nxt_idx, nxt);

        // ---- interpolation to the quadrature points, all in registers: u[i][j][k] -> v[p][q][r]
        double v[N3];
        if constexpr (COLL) {
#pragma unroll
            for (int l = 0; l < N3; ++l) v[l] = cur[l];  // (a register copy the compiler folds away)
        } else {
            double t1[NM * NM * NQ], t2[NM * NQ * NQ];
#pragma unroll
            for (int ij = 0; ij < NM * NM; ++ij)
#pragma unroll
                for (int r = 0; r < NQ; ++r) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < NM; ++k) s = fma(m.B[r * NM + k], cur[ij * NM + k], s);
                    t1[ij * NQ + r] = s;
                }
#pragma unroll
            for (int i = 0; i < NM; ++i)
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int r = 0; r < NQ; ++r) {
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < NM; ++j) s = fma(m.B[q * NM + j], t1[(i * NM + j) * NQ + r], s);
                        t2[(i * NQ + q) * NQ + r] = s;
                    }
#pragma unroll
            for (int p = 0; p < NQ; ++p)
#pragma unroll
                for (int qr = 0; qr < NQ * NQ; ++qr) {
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i < NM; ++i) s = fma(m.B[p * NM + i], t2[i * NQ * NQ + qr], s);
                    v[p * NQ * NQ + qr] = s;
                }
        }

        // ---- operator at the points: w = D^T G D v (+ JxW v)
        double w[N3];
#pragma unroll
        for (int l = 0; l < N3; ++l) w[l] = 0.0;
        if constexpr (STAGED) {
            v2::mbar_wait(bar, parity);  // this batch's geometric factors / JxW have landed
            parity ^= 1u;
        }
        if constexpr (LAP) {
#pragma unroll
            for (int p = 0; p < NQ; ++p)
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int r = 0; r < NQ; ++r) {
                        const int pt = (p * NQ + q) * NQ + r;
                        double gr = 0.0, gs = 0.0, gt = 0.0;
#pragma unroll
                        for (int n = 0; n < NQ; ++n) {
                            gr = fma(m.D[p * NQ + n], v[(n * NQ + q) * NQ + r], gr);
                            gs = fma(m.D[q * NQ + n], v[(p * NQ + n) * NQ + r], gs);
                            gt = fma(m.D[r * NQ + n], v[(p * NQ + q) * NQ + n], gt);
                        }
                        const double g0 = Ge[0 * N3 + pt], g1 = Ge[1 * N3 + pt], g2 = Ge[2 * N3 + pt];
                        const double g3 = Ge[3 * N3 + pt], g4 = Ge[4 * N3 + pt], g5 = Ge[5 * N3 + pt];
                        const double fr = g0 * gr + g1 * gs + g2 * gt;
                        const double fs = g1 * gr + g3 * gs + g4 * gt;
                        const double ft = g2 * gr + g4 * gs + g5 * gt;
                        if constexpr (LVEC) dot_acc = fma(gr, fr, fma(gs, fs, fma(gt, ft, dot_acc)));
#pragma unroll
                        for (int n = 0; n < NQ; ++n) {
                            w[(n * NQ + q) * NQ + r] = fma(m.D[p * NQ + n], fr, w[(n * NQ + q) * NQ + r]);
                            w[(p * NQ + n) * NQ + r] = fma(m.D[q * NQ + n], fs, w[(p * NQ + n) * NQ + r]);
                            w[(p * NQ + q) * NQ + n] = fma(m.D[r * NQ + n], ft, w[(p * NQ + q) * NQ + n]);
                        }
                    }
        }
        if constexpr (MASS) {
#pragma unroll
            for (int l = 0; l < N3; ++l) {
                const double mv = (active ? Je[l] : 0.0) * v[l];
                w[l] += mv;
                if constexpr (LVEC) dot_acc = fma(mv, v[l], dot_acc);
            }
        }
        if constexpr (STAGED) {
            __syncthreads();  // every thread has drained its slots: the buffers may be refilled
            if (nb < n_batches) {
                v2::fence_proxy_async();
                issue_g(nb);
            }
        }
        if constexpr (LVEC && PREFETCH_VAL) load_val(nb, nxt_idx, nxt);  // next batch's gathers behind the rest of this element

        // ---- back to the nodes and out
        double z[M3];
        if constexpr (COLL) {
#pragma unroll
            for (int l = 0; l < N3; ++l) z[l] = w[l];
        } else {
            double t2[NM * NQ * NQ], t1[NM * NM * NQ];
#pragma unroll
            for (int i = 0; i < NM; ++i)
#pragma unroll
                for (int qr = 0; qr < NQ * NQ; ++qr) {
                    double s = 0.0;
#pragma unroll
                    for (int p = 0; p < NQ; ++p) s = fma(m.B[p * NM + i], w[p * NQ * NQ + qr], s);
                    t2[i * NQ * NQ + qr] = s;
                }
#pragma unroll
            for (int i = 0; i < NM; ++i)
#pragma unroll
                for (int j = 0; j < NM; ++j)
#pragma unroll
                    for (int r = 0; r < NQ; ++r) {
                        double s = 0.0;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) s = fma(m.B[q * NM + j], t2[(i * NQ + q) * NQ + r], s);
                        t1[(i * NM + j) * NQ + r] = s;
                    }
#pragma unroll
            for (int ij = 0; ij < NM * NM; ++ij)
#pragma unroll
                for (int k = 0; k < NM; ++k) {
                    double s = 0.0;
#pragma unroll
                    for (int r = 0; r < NQ; ++r) s = fma(m.B[r * NM + k], t1[ij * NQ + r], s);
                    z[ij * NM + k] = s;
                }
        }
        if constexpr (LVEC) {
#pragma unroll
            for (int l = 0; l < M3; ++l)
                if (cur_idx[l] != kInvalidIndex) atomicAdd(a.out + cur_idx[l], z[l]);
        } else if (active) {
#pragma unroll
            for (int l = 0; l < M3; ++l) a.out[(size_t)e * M3 + l] = z[l];
        }
        if constexpr (LVEC) {
#pragma unroll
            for (int l = 0; l < M3; ++l) cur_idx[l] = nxt_idx[l];
        }
        if constexpr (PREFETCH_VAL) {
#pragma unroll
            for (int l = 0; l < M3; ++l) cur[l] = nxt[l];
        } else
            load_val(nb, cur_idx, cur);
    }

    if constexpr (LVEC) {
        if (a.dot != nullptr) {
            for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_xor_sync(0xffffffffu, dot_acc, o);  // TPB is a multiple of 32
            if ((tid & 31) == 0 && dot_acc != 0.0) atomicAdd(a.dot, dot_acc);
        }
    }
}

}  // namespace b200fe
