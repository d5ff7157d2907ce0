// sumfact_mass.cuh -- mass-operator cell kernel (CEED BK1 / BP1: out = B^T (JxW .* (B u)), nq = p + 2) with ONE shared-memory
// round trip per direction instead of three.
//
// Why (ncu, profiles/r01d_bk_l1tex.txt; bench r02m): the generic kernel runs every 1-D sweep as "register column x matrix"
// and changes the layout through shared memory between sweeps -- six round trips for the interpolated mass operator, whose
// only DRAM traffic is 8 (2 nm^3 + nq^3) bytes per element.  Its shared-memory traffic is four times its DRAM traffic and
// half of the wavefronts are partial (nm^2 or nm nq active threads of nq^2), so BK1 / BP1 sit at 0.5 of the HBM roofline with
// L1TEX at 87-89 %.  Here a thread owns a whole PLANE of an element in registers for the two outer directions:
//   phase A  thread (element, i):   u[i][:][:]  --k->r, j->q in registers-->  t[i][:][:]  (nq^2 values)        -> shared memory
//   phase B  thread (element, q, r): column t[:][q][r] --i->p--> v[p];  w = JxW v;  --p->i--> s[:][q][r]        -> shared memory
//   phase C  thread (element, i):   s[i][:][:]  --q->j, r->k in registers-->  out[i][:][:]
// 4 nm nq^2 shared-memory accesses per element instead of ~ 15 nm^2 nq, all of them full, conflict-free wavefronts (odd plane
// stride); phases A and C belong to the same thread, so one barrier on each side of phase B is all the synchronisation.
// The element vectors are read / gathered and written / scattered by the plane threads (nm^2 consecutive values each);
// JxW is streamed by the phase-B threads, coalesced, one batch ahead.  Fused p.Ap = sum JxW v^2 (L-vector mode).
// 1-D products go through the same interp / interp_t wrappers as sumfact2.cuh (plain or even-odd contractions).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sumfact2.cuh"

namespace b200fe {

template <int NM, int NQ, int EPB>
struct MassLayout {
    static constexpr int N2 = NQ * NQ, N3 = N2 * NQ, M2 = NM * NM, M3 = M2 * NM;
    static constexpr int PS = N2 | 1;                       // odd plane stride: thread a writes / reads plane a*PS .. conflict-free
    static constexpr int THREADS = ((EPB * N2 + 31) / 32) * 32;
    static constexpr int PLANE_THREADS = EPB * NM;          // the first EPB*NM threads also act as plane threads
    static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)EPB * NM * PS; }
};

template <int NM, int NQ, bool LVEC, int EPB, int MINB, bool EO>
__global__ void __launch_bounds__((MassLayout<NM, NQ, EPB>::THREADS), MINB)
    sumfact_mass_kernel(const __grid_constant__ Mats<NM, NQ, EO> m, const KArgs a)
{
    using L = MassLayout<NM, NQ, EPB>;
    constexpr int N2 = L::N2, N3 = L::N3, M2 = L::M2, M3 = L::M3, PS = L::PS;
    if (a.skip != nullptr && *a.skip != 0) return;
    extern __shared__ __align__(16) double S[];  // [EPB][NM][PS]
    const int tid = threadIdx.x;
    // plane role: (element slot ea, plane i); column role: (element slot eb, point qr)
    const bool plane_thread = tid < L::PLANE_THREADS;
    const int ea = tid / NM, pi = tid - ea * NM;
    const bool col_thread = tid < EPB * N2;
    const int eb_ = tid / N2, qr = tid - eb_ * N2;
    double *Sp = S + (size_t)tid * PS;                          // my plane (plane role): (ea*NM + pi)*PS = tid*PS
    double *Sc = S + (size_t)(col_thread ? eb_ : 0) * NM * PS + qr;   // my column (column role): + i*PS
    const uint32_t n_batches = (a.n_elems + EPB - 1) / EPB;

    // JxW column of the column role, one batch ahead
    double jw_next[NQ];
    auto load_jw = [&](uint32_t batch, double (&j)[NQ]) {
        const uint32_t e_ = batch * EPB + eb_;
        const bool ok = col_thread && batch < n_batches && e_ < a.n_elems;
#pragma unroll
        for (int p = 0; p < NQ; ++p) j[p] = ok ? __ldg(a.JxW + (size_t)e_ * N3 + p * N2 + qr) : 0.0;
    };
    load_jw(blockIdx.x, jw_next);

    double dot_acc = 0.0;
    for (uint32_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        const uint32_t e_plane = batch * EPB + ea;
        const bool plane_ok = plane_thread && e_plane < a.n_elems;
        [[maybe_unused]] uint32_t idx[M2];
        // ---- phase A: my plane u[i][j][k] -> t[i][q][r]
        if (plane_thread) {
            double u[NM][NM];
            if constexpr (LVEC) {
#pragma unroll
                for (int l = 0; l < M2; ++l) idx[l] = plane_ok ? __ldg(a.idx + (size_t)e_plane * M3 + pi * M2 + l) : kInvalidIndex;
#pragma unroll
                for (int l = 0; l < M2; ++l) u[l / NM][l % NM] = idx[l] == kInvalidIndex ? 0.0 : __ldg(a.in + idx[l]);
            } else {
#pragma unroll
                for (int l = 0; l < M2; ++l) u[l / NM][l % NM] = plane_ok ? __ldg(a.in + (size_t)e_plane * M3 + pi * M2 + l) : 0.0;
            }
            double t1[NM][NQ];  // [j][r]
#pragma unroll
            for (int j = 0; j < NM; ++j) v2::interp<NM, NQ>(m, u[j], t1[j]);
#pragma unroll
            for (int r = 0; r < NQ; ++r) {  // j -> q for fixed r, straight into shared memory
                double col[NM], out[NQ];
#pragma unroll
                for (int j = 0; j < NM; ++j) col[j] = t1[j][r];
                v2::interp<NM, NQ>(m, col, out);
#pragma unroll
                for (int q = 0; q < NQ; ++q) Sp[q * NQ + r] = out[q];
            }
        }
        __syncthreads();
        // ---- phase B: my column over i -> points -> JxW -> back over i, in place
        if (col_thread) {
            double jw[NQ];
#pragma unroll
            for (int p = 0; p < NQ; ++p) jw[p] = jw_next[p];
            double c[NM], v[NQ];
#pragma unroll
            for (int i = 0; i < NM; ++i) c[i] = Sc[i * PS];
            v2::interp<NM, NQ>(m, c, v);
#pragma unroll
            for (int p = 0; p < NQ; ++p) {
                const double mv = jw[p] * v[p];
                if constexpr (LVEC) dot_acc = fma(mv, v[p], dot_acc);  // (jw = 0 on idle slots)
                v[p] = mv;
            }
            v2::interp_t<NM, NQ>(m, v, c);
#pragma unroll
            for (int i = 0; i < NM; ++i) Sc[i * PS] = c[i];
        }
        load_jw(batch + gridDim.x, jw_next);  // next batch's JxW behind phase C
        __syncthreads();
        // ---- phase C: my plane s[i][q][r] -> out[i][j][k]
        if (plane_thread) {
            double t1[NM][NQ];  // [j][r]
#pragma unroll
            for (int r = 0; r < NQ; ++r) {
                double col[NQ], out[NM];
#pragma unroll
                for (int q = 0; q < NQ; ++q) col[q] = Sp[q * NQ + r];
                v2::interp_t<NM, NQ>(m, col, out);
#pragma unroll
                for (int j = 0; j < NM; ++j) t1[j][r] = out[j];
            }
#pragma unroll
            for (int j = 0; j < NM; ++j) {
                double z[NM];
                v2::interp_t<NM, NQ>(m, t1[j], z);
                if constexpr (LVEC) {
#pragma unroll
                    for (int k = 0; k < NM; ++k)
                        if (idx[j * NM + k] != kInvalidIndex) atomicAdd(a.out + idx[j * NM + k], z[k]);
                } else if (plane_ok) {
#pragma unroll
                    for (int k = 0; k < NM; ++k) a.out[(size_t)e_plane * M3 + pi * M2 + j * NM + k] = z[k];
                }
            }
        }
        // (the next batch's phase A writes the planes these same threads have just read: no barrier needed here)
    }
    if constexpr (LVEC) {
        if (a.dot != nullptr) {
            for (int o = 16; o > 0; o >>= 1) dot_acc += __shfl_xor_sync(0xffffffffu, dot_acc, o);  // THREADS is a multiple of 32
            if ((tid & 31) == 0 && dot_acc != 0.0) atomicAdd(a.dot, dot_acc);
        }
    }
}

}  // namespace b200fe
