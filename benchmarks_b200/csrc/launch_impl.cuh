// launch_impl.cuh -- launcher template body; included only by inst.cu.
#pragma once
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "sumfact2.cuh"

namespace b200fe {

// Launch-shape heuristics (round-1 ncu sweep, profiles/r01_bk_variants.md): small planes want
// ~128-thread CTAs, large planes want two CTAs per SM (register cap 128).  -D overrides are
// tuning knobs for tools/gpu_run*.sh only.
constexpr int tpb_for(int nq)
{
#ifdef B200FE_TPB
    return B200FE_TPB;
#else
    return nq <= 6 ? 128 : 256;
#endif
}
constexpr int minb_for(int nq)
{
#ifdef B200FE_MINB
    return B200FE_MINB;
#else
    return nq >= 6 ? 2 : 1;
#endif
}

// elements per CTA: fill ~tpb_for(nq) threads with whole quadrature planes
constexpr int epb_for(int nq)
{
    const int n2 = nq * nq;
    int e = tpb_for(nq) / n2;
    return e < 1 ? 1 : e;
}

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
constexpr int v2_rmin(int nq, bool coll, int qop)
{
#ifdef B200FE_V2_RMIN_FIXED
    return B200FE_V2_RMIN_FIXED;
#else
    if (!(qop & QOP_LAPLACE)) return 64 + 12 * nq;  // mass only: no G buffer, registers are the limit
    if (nq <= 6) return (qop & QOP_AFFINE) ? 128 : 96;  // affine kernels keep six more constants + weights live
#ifdef B200FE_V2_RMIN_HI
    return B200FE_V2_RMIN_HI;
#else
    if (nq == 7) return coll ? 160 : 240;
    if (nq == 8) return coll ? 208 : 240;
    return coll ? 208 : 160;
#endif
#endif
}

template <int NM, int NQ, bool COLL, int QOP>
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
    static constexpr int BY_REGS = 65536 / (v2_rmin(NQ, COLL, QOP) * T32);
    static constexpr int BY_THREADS = 2048 / T32;
    static constexpr int M0 = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
    static constexpr int M1 = M0 < BY_THREADS ? M0 : BY_THREADS;
#ifdef B200FE_MINB
    static constexpr int MINB = B200FE_MINB;
#else
    static constexpr int MINB = M1 < 1 ? 1 : (M1 > 16 ? 16 : M1);
#endif
};

template <int NM, int NQ, bool COLL, int QOP, bool LVEC>
cudaError_t launch_t(const double *hB, const double *hD, const double *hW, const KArgs &a, cudaStream_t s,
                     LaunchInfo *info, bool dry_run)
{
#ifdef B200FE_KERNEL_V1
    constexpr int EPB = epb_for(NQ);
    constexpr int T = EPB * NQ * NQ;
    using L = Layout<NM, NQ, COLL>;
    auto kern = sumfact_kernel<NM, NQ, COLL, QOP, LVEC, EPB, minb_for(NQ)>;
    const size_t smem = L::smem_bytes(EPB);
#else
    using C = V2Cfg<NM, NQ, COLL, QOP>;
    constexpr int EPB = C::EPB;
    constexpr int T = C::T;
    auto kern = sumfact2_kernel<NM, NQ, COLL, QOP, LVEC, EPB, C::MINB>;
    const size_t smem = C::SMEM;
    // TMA bulk copies need a 16-byte aligned source (the batch block offset is a multiple of 48 nq^3 bytes)
    if ((QOP & QOP_LAPLACE) && !(QOP & QOP_AFFINE) && !dry_run && (reinterpret_cast<uintptr_t>(a.G) & 15u) != 0) return cudaErrorMisalignedAddress;
#endif

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
    if (info) *info = LaunchInfo{EPB, grid, T, (int)smem, c.blocks_per_sm, c.regs};
    if (dry_run || a.n_elems == 0) return cudaSuccess;

    Mats<NM, NQ> m;
    if (hB) std::memcpy(m.B, hB, sizeof(m.B)); else std::memset(m.B, 0, sizeof(m.B));
    if (hD) std::memcpy(m.D, hD, sizeof(m.D)); else std::memset(m.D, 0, sizeof(m.D));
    if (hW) std::memcpy(m.W, hW, sizeof(m.W)); else std::memset(m.W, 0, sizeof(m.W));
#ifdef B200FE_EVEN_ODD
    // tuning variant: only the symmetric matrices of a real basis (not the reference's cos() test matrices)
    if (eo::fill<NM, NQ>(m.B, m.D, m.E) > 1e-10) return cudaErrorNotSupported;
#endif
    kern<<<grid, T, smem, s>>>(m, a);
    return cudaGetLastError();
}

}  // namespace b200fe
