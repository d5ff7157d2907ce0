// launch_impl.cuh -- launcher template body; included only by inst.cu.
#pragma once
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "sumfact2.cuh"
#include "sumfact_tpe.cuh"
#include "sumfact_cart.cuh"

namespace b200fe {

// v2 kernel: target threads per CTA.  r01 sweeps (profiles/r01_v2_variants.txt): small planes want
// several elements per CTA (~160 threads), nq >= 7 wants one element per CTA.
#ifdef B200FE_V2_TPB
#define B200FE_V2_TPB_FOR(nq) (B200FE_V2_TPB)
#else
#define B200FE_V2_TPB_FOR(nq) ((nq) <= 6 ? (COLL ? 160 : 128) : 96)  // r01 sweep: bk3 p=3,4 71/70 % at 128 vs 68/67 % at 160
#endif
// v2 kernel: registers per thread that the occupancy target (MINB) must leave.  Data-driven
// (profiles/r01_v2_variants.txt): with the software-pipelined inputs the large planes need 160-240
// registers to stay spill-free; below nq = 7 a 96-register floor (more CTAs per SM) wins.
constexpr int v2_rmin(int nq, bool coll, int qop, bool eo = false)
{
    // Register caps come in steps: the register file is split over 4 SM sub-partitions, so a CTA of 3 warps is capped at
    // 168 registers with 3 or 4 resident CTAs and at 255 with 2 (ptxas honours exactly that).  Measured A/B (profiles/r02f_*):
    //  * even-odd interpolated Laplace at nq = 10: 2 CTAs x 255 registers, no spill (was 3 x 168 with 104 B/thread of local
    //    memory): BP3 p = 8 0.611 -> 0.649 of the HBM roofline;
    //  * mass at nq = 9 (plain contractions): 2 x 255 instead of 3 x 168 + 208 B spill: BK1 p = 7 55.9 -> 61.9 GDoF/s.
#ifndef B200FE_V2_RMIN_EO10
#define B200FE_V2_RMIN_EO10 255
#endif
#ifndef B200FE_V2_RMIN_MASS9
#define B200FE_V2_RMIN_MASS9 255
#endif
#ifdef B200FE_V2_RMIN_EO9
    if (eo && !coll && nq == 9 && (qop & QOP_LAPLACE)) return B200FE_V2_RMIN_EO9;
#endif
    // plain contractions at nq = 9, 10 (only reached with non-symmetric 1-D matrices, i.e. the reference drivers' cos() test
    // matrices on the E-vector kernels): tuning knobs for the same 2 x 255 vs 3-4 x 168 question
#ifdef B200FE_V2_RMIN_GEN9
    if (!eo && !coll && nq == 9 && (qop & QOP_LAPLACE)) return B200FE_V2_RMIN_GEN9;
#endif
#ifdef B200FE_V2_RMIN_GEN10
    if (!eo && !coll && nq == 10 && (qop & QOP_LAPLACE)) return B200FE_V2_RMIN_GEN10;
#endif
    if (eo && !coll && nq == 10 && (qop & QOP_LAPLACE)) return B200FE_V2_RMIN_EO10;
    if (!(qop & QOP_LAPLACE) && nq == 9) return B200FE_V2_RMIN_MASS9;
#ifdef B200FE_V2_RMIN_FIXED
    return B200FE_V2_RMIN_FIXED;
#else
    if (!(qop & QOP_LAPLACE)) return 64 + 12 * nq;  // mass only: no G buffer, registers are the limit
    // separable kernel for axis-aligned cells: no G stream to wait for, three contractions -- the even-odd build needs 108
    // registers at nq = 7 and 167 at nq = 9; more resident CTAs hide the gather latency
#ifdef B200FE_CART_COLL_RMIN
    if (qop & QOP_CARTESIAN) return B200FE_CART_COLL_RMIN;  // tuning knob
#endif
    if (qop & QOP_CARTESIAN) return nq <= 6 ? 96 : nq == 7 ? 112 : nq == 8 ? 128 : 168;
    if (nq <= 6) return (qop & (QOP_AFFINE | QOP_TRILINEAR)) ? 128 : 96;  // on-the-fly kernels keep the cell constants + weights live
#ifdef B200FE_V2_RMIN_HI
    return B200FE_V2_RMIN_HI;
#else
    if (nq == 7) return coll ? 160 : 240;
    if (nq == 8) return coll ? 208 : 240;
    return coll ? 208 : 160;
#endif
#endif
}

template <int NM, int NQ, bool COLL, int QOP, bool EO = false>
struct V2Cfg {
    using L = v2::Layout2<NM, NQ, COLL, QOP>;
#ifndef B200FE_V2_WL_WARPS
#define B200FE_V2_WL_WARPS 4   // warp-local mode (nq <= 5): warps per CTA
#endif
    static constexpr int EPB_PLAIN = (B200FE_V2_TPB_FOR(NQ) / (NQ * NQ)) < 1 ? 1 : (B200FE_V2_TPB_FOR(NQ) / (NQ * NQ));
    static constexpr int EPB = L::WARP_LOCAL ? B200FE_V2_WL_WARPS * L::EPW : EPB_PLAIN;
    static constexpr int T = L::threads(EPB);
    static constexpr int T32 = (T + 31) / 32 * 32;
    static constexpr size_t SMEM = L::smem_bytes(EPB);
    static constexpr int BY_SMEM = (int)((227 * 1024) / (SMEM + 1024));
    static constexpr int BY_REGS = 65536 / (v2_rmin(NQ, COLL, QOP, EO) * T32);
    static constexpr int BY_THREADS = 2048 / T32;
    static constexpr int M0 = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
    static constexpr int M1 = M0 < BY_THREADS ? M0 : BY_THREADS;
#ifdef B200FE_MINB
    static constexpr int MINB = B200FE_MINB;
#else
    static constexpr int MINB = M1 < 1 ? 1 : (M1 > 16 ? 16 : M1);
#endif
};

// Even-odd kernels: built where the B200 sweeps show a gain (profiles/r02a_sweep_{default,eo}.txt): every interpolated
// operator, and the collocated ones from nq = 9 on (below that the collocated kernels are at 0.87-0.93 of the HBM roofline
// either way and the plain contraction is as fast or faster: BP5 p = 7 0.92 vs 0.87).
// The on-the-fly (affine) geometry kernels keep six per-cell constants and the weights live on top of the contractions and
// spilled 424-1208 B/thread with plain contractions (profiles/r01k_static_resource_usage.txt): always even-odd.
constexpr bool eo_built(int nq, bool coll, int qop = 0) { return (qop & (QOP_AFFINE | QOP_TRILINEAR)) ? true : (coll ? nq >= 9 : true); }

inline bool eo_enabled()
{
    static const bool on = [] { const char *e = std::getenv("B200FE_EVEN_ODD"); return !e || std::atoi(e) != 0; }();
    return on;
}

// multi-component kernels (G streamed once for all components of a vector-valued operator): collocated Laplace L-vector
constexpr bool mc_built(bool coll, int qop, bool lvec) { return coll && lvec && qop == QOP_LAPLACE; }

template <int NM, int NQ, bool COLL, int QOP, bool LVEC, bool EO, bool MC = false>
cudaError_t launch_variant(const Mats<NM, NQ, EO> &m, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry_run)
{
    using C = V2Cfg<NM, NQ, COLL, QOP, EO>;
    constexpr int EPB = C::EPB;
    constexpr int T = C::T;
    auto kern = sumfact2_kernel<NM, NQ, COLL, QOP, LVEC, EPB, C::MINB, EO, MC>;
    const size_t smem = C::SMEM;
    // TMA bulk copies need a 16-byte aligned source (the batch block offset is a multiple of 48 nq^3 bytes)
    if ((QOP & QOP_LAPLACE) && !(QOP & (QOP_AFFINE | QOP_TRILINEAR)) && !dry_run && (reinterpret_cast<uintptr_t>(a.G) & 15u) != 0) return cudaErrorMisalignedAddress;

    struct Cfg {
        bool ready = false;
        int blocks_per_sm = 0, sms = 0, regs = 0;
    };
    static Cfg cfg[64];  // per device
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    Cfg &c = cfg[dev & 63];
    if (!c.ready) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.blocks_per_sm, kern, T, smem);
        if (err != cudaSuccess) return err;
        err = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return err;
        cudaFuncAttributes fa;
        err = cudaFuncGetAttributes(&fa, kern);
        if (err != cudaSuccess) return err;
        c.regs = fa.numRegs;
        if (c.blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
        c.ready = true;
    }
    const uint32_t n_batches = (a.n_elems + EPB - 1) / EPB;
    long long resident = (long long)c.sms * c.blocks_per_sm * grid_multiplier();
    const int grid = (int)(n_batches < (uint32_t)resident ? n_batches : resident);
    if (info) *info = LaunchInfo{EPB, grid, T, (int)smem, c.blocks_per_sm, c.regs, EO ? 1 : 0};
    if (dry_run || a.n_elems == 0) return cudaSuccess;
    kern<<<grid, T, smem, s>>>(m, a);
    return cudaGetLastError();
}

// tiny elements (nq <= 3): one thread per element, everything in registers (sumfact_tpe.cuh); B200FE_TPE=0 keeps the
// plane-per-thread kernel for A/B runs.  Measured (profiles/r02k_tpe_ab.txt, fraction of the HBM roofline, plane-per-thread
// -> thread-per-element): BP5 p=1 0.61 -> 0.95, BP3 p=1 0.60 -> 0.93, BK3 p=1 0.58 -> 0.92, BK5 p=1 0.65 -> 0.83, BP5 p=2
// 0.71 -> 0.75, "bp35" p=2 0.58 -> 0.74.  Operators with geometric factors only: the pure mass kernels read JxW straight
// from global memory per thread and lose (BK1 p=1 0.42 -> 0.36), they keep the plane-per-thread kernel.
constexpr bool tpe_built(int nq, int qop) { return nq <= 3 && !(qop & (QOP_AFFINE | QOP_TRILINEAR)); }
inline bool tpe_enabled()
{
    static const bool on = [] { const char *e = std::getenv("B200FE_TPE"); return !e || std::atoi(e) != 0; }();
    return on;
}

template <int NM, int NQ, bool COLL, int QOP, bool LVEC>
cudaError_t launch_tpe(const double *hB, const double *hD, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry_run)
{
    constexpr int TPB = NQ == 2 ? 64 : 32;
    constexpr int MINB = NQ == 2 ? 6 : 1;
    using L = tpe::LayoutT<NM, NQ, COLL, QOP, TPB>;
    auto kern = sumfact_tpe_kernel<NM, NQ, COLL, QOP, LVEC, TPB, MINB>;
    const size_t smem = L::smem_bytes();
    if ((QOP & QOP_LAPLACE) && !dry_run && (reinterpret_cast<uintptr_t>(a.G) & 15u) != 0) return cudaErrorMisalignedAddress;
    struct Cfg {
        bool ready = false;
        int blocks_per_sm = 0, sms = 0, regs = 0;
    };
    static Cfg cfg[64];
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    Cfg &c = cfg[dev & 63];
    if (!c.ready) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.blocks_per_sm, kern, TPB, smem);
        if (err != cudaSuccess) return err;
        err = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return err;
        cudaFuncAttributes fa;
        err = cudaFuncGetAttributes(&fa, kern);
        if (err != cudaSuccess) return err;
        c.regs = fa.numRegs;
        if (c.blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
        c.ready = true;
    }
    const uint32_t n_batches = (a.n_elems + TPB - 1) / TPB;
    const long long resident = (long long)c.sms * c.blocks_per_sm * grid_multiplier();
    const int grid = (int)(n_batches < (uint32_t)resident ? n_batches : resident);
    if (info) *info = LaunchInfo{TPB, grid, TPB, (int)smem, c.blocks_per_sm, c.regs, 0};
    if (dry_run || a.n_elems == 0) return cudaSuccess;
    Mats<NM, NQ, false> m;
    if (hB) std::memcpy(m.B, hB, sizeof(m.B));
    else for (int q = 0; q < NQ; ++q) for (int i = 0; i < NM; ++i) m.B[q * NM + i] = (COLL && q == i) ? 1.0 : 0.0;
    if (hD) std::memcpy(m.D, hD, sizeof(m.D)); else std::memset(m.D, 0, sizeof(m.D));
    std::memset(m.W, 0, sizeof(m.W));
    std::memset(m.X, 0, sizeof(m.X));
    kern<<<grid, TPB, smem, s>>>(m, a);
    return cudaGetLastError();
}

template <int NM, int QOP>
cudaError_t launch_cart_q(const double *hKM, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry_run)
{
    using L = cart::LayoutC<NM>;
    constexpr int N2 = NM * NM;
#ifndef B200FE_CART_TPB
#define B200FE_CART_TPB 128  // target threads per CTA (tuning knob)
#endif
    constexpr int EPB = B200FE_CART_TPB / N2 < 1 ? 1 : B200FE_CART_TPB / N2;
    constexpr int T = EPB * N2, T32 = (T + 31) / 32 * 32;
#ifdef B200FE_CART_RMIN
    constexpr int RMIN = B200FE_CART_RMIN;  // tuning knob: one register floor for every degree
#else
    // registers the occupancy target must leave: what the packed even-odd build needs without spilling (ptxas: 89 / 103 /
    // 116 / 149 / 198 registers at nm = 5 ... 9; the first build with full matrices wanted > 255 at nm = 7, 9)
    constexpr int RMIN = NM <= 4 ? 72 + 10 * NM : NM == 5 ? 96 : NM == 6 ? 104 : NM == 7 ? 120 : NM == 8 ? 152 : 168;  // (nm = 9: 4 CTAs of 3 warps cap at 168, 68 B/thread spilled)
#endif
    constexpr int BY_REGS = 65536 / (RMIN * T32), BY_THREADS = 2048 / T32;
    constexpr int M0 = BY_REGS < BY_THREADS ? BY_REGS : BY_THREADS;
    constexpr int MINB = M0 < 1 ? 1 : (M0 > 16 ? 16 : M0);
    auto kern = sumfact_cart_kernel<NM, EPB, MINB, QOP>;
    const size_t smem = L::smem_bytes(EPB);
    struct Cfg {
        bool ready = false;
        int blocks_per_sm = 0, sms = 0, regs = 0;
    };
    static Cfg cfg[64];  // per device
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    Cfg &c = cfg[dev & 63];
    if (!c.ready) {
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (err != cudaSuccess) return err;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.blocks_per_sm, kern, T, smem);
        if (err != cudaSuccess) return err;
        err = cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, dev);
        if (err != cudaSuccess) return err;
        cudaFuncAttributes fa;
        err = cudaFuncGetAttributes(&fa, kern);
        if (err != cudaSuccess) return err;
        c.regs = fa.numRegs;
        if (c.blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
        c.ready = true;
    }
    const uint32_t n_batches = (a.n_elems + EPB - 1) / EPB;
    const long long resident = (long long)c.sms * c.blocks_per_sm * grid_multiplier();
    const int grid = (int)(n_batches < (uint32_t)resident ? n_batches : resident);
    if (info) *info = LaunchInfo{EPB, grid, T, (int)smem, c.blocks_per_sm, c.regs, 0};
    if (dry_run || a.n_elems == 0) return cudaSuccess;
    if (hKM == nullptr) return cudaErrorInvalidValue;
    CartMats<NM> m;
    // (b200fe_op_create only selects this kernel for symmetric, point-symmetric K and M; refuse anything else)
    if (m.K.fill(hKM) > 1e-10 || m.M.fill(hKM + NM * NM) > 1e-10) return cudaErrorInvalidValue;
    kern<<<grid, T, smem, s>>>(m, a);
    return cudaGetLastError();
}

template <int NM>
cudaError_t launch_cart_t(int qop, const double *hKM, const KArgs &a, cudaStream_t s, LaunchInfo *info, bool dry_run)
{
    if (qop == QOP_LAPLACE) return launch_cart_q<NM, QOP_LAPLACE>(hKM, a, s, info, dry_run);
    if (qop == QOP_MASS) return launch_cart_q<NM, QOP_MASS>(hKM, a, s, info, dry_run);
    if (qop == QOP_HELMHOLTZ) return launch_cart_q<NM, QOP_HELMHOLTZ>(hKM, a, s, info, dry_run);
    return cudaErrorInvalidValue;
}

template <int NM, int NQ, bool COLL, int QOP, bool LVEC>
cudaError_t launch_t(const double *hB, const double *hD, const double *hW, const KArgs &a, cudaStream_t s,
                     LaunchInfo *info, bool dry_run)
{
    if constexpr (tpe_built(NQ, QOP)) {
        // (the staged JxW batch copy needs a 16-byte aligned source: a cell range that starts at an odd cell of an nq = 3
        //  mass-type operator falls back to the plane-per-thread kernel)
        const bool jxw_ok = !(QOP & QOP_MASS) || dry_run || (reinterpret_cast<uintptr_t>(a.JxW) & 15u) == 0;
        if (tpe_enabled() && a.ncomp <= 1 && jxw_ok) return launch_tpe<NM, NQ, COLL, QOP, LVEC>(hB, hD, a, s, info, dry_run);
    }
    if constexpr (eo_built(NQ, COLL, QOP)) {
        if (eo_enabled()) {
            // the symmetric 1-D matrices of a real basis (not the reference drivers' cos() test matrices): even-odd kernel
            double Bsym[NQ * NM], Dsym[NQ * NQ];
            if (hB) std::memcpy(Bsym, hB, sizeof(Bsym));
            else for (int q = 0; q < NQ; ++q) for (int i = 0; i < NM; ++i) Bsym[q * NM + i] = (COLL && q == i) ? 1.0 : 0.0;
            if (hD) std::memcpy(Dsym, hD, sizeof(Dsym)); else std::memset(Dsym, 0, sizeof(Dsym));
            Mats<NM, NQ, true> me;
            if (eo::fill<NM, NQ>(Bsym, Dsym, me.E) <= 1e-10) {
                if (hW) std::memcpy(me.W, hW, sizeof(me.W)); else std::memset(me.W, 0, sizeof(me.W));
                if (hW && (QOP & QOP_TRILINEAR)) std::memcpy(me.X, hW + NQ, sizeof(me.X)); else std::memset(me.X, 0, sizeof(me.X));  // hW = weights | points
                if (a.ncomp > 1) {
                    if constexpr (mc_built(COLL, QOP, LVEC)) return launch_variant<NM, NQ, COLL, QOP, LVEC, true, true>(me, a, s, info, dry_run);
                    else return cudaErrorNotSupported;
                }
                return launch_variant<NM, NQ, COLL, QOP, LVEC, true>(me, a, s, info, dry_run);
            }
        }
    }
    if constexpr ((QOP & QOP_TRILINEAR) != 0) {
        return cudaErrorNotSupported;  // trilinear on-the-fly geometry is built with the even-odd contractions only (real bases)
    } else {
    Mats<NM, NQ, false> m;
    if (hB) std::memcpy(m.B, hB, sizeof(m.B)); else std::memset(m.B, 0, sizeof(m.B));
    if (hD) std::memcpy(m.D, hD, sizeof(m.D)); else std::memset(m.D, 0, sizeof(m.D));
    if (hW) std::memcpy(m.W, hW, sizeof(m.W)); else std::memset(m.W, 0, sizeof(m.W));
    std::memset(m.X, 0, sizeof(m.X));
    if (a.ncomp > 1) {
        if constexpr (mc_built(COLL, QOP, LVEC)) return launch_variant<NM, NQ, COLL, QOP, LVEC, false, true>(m, a, s, info, dry_run);
        else return cudaErrorNotSupported;
    }
    return launch_variant<NM, NQ, COLL, QOP, LVEC, false>(m, a, s, info, dry_run);
    }
}

}  // namespace b200fe
