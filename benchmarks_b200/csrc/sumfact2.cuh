// sumfact2.cuh -- the cell kernel for sm_100a (FP64): TMA-staged geometric factors + register-column contractions in
// all three directions.
//
// Why (ncu, profiles/r01a_bp5_p6_kernel_summary.txt): the first-generation kernel (removed in round 2) was bound by
// exposed HBM latency (stall long_scoreboard 10.6 per issue at 25 % occupancy) and by shared-memory
// wavefronts (881 per element at nq = 7, equal to the whole HBM-roofline cycle budget).  Changes:
//
//  * G (6 nq^3 doubles per element = 75 % of all traffic) no longer goes through registers: the
//    CTA's whole element batch is one contiguous 16-byte-aligned block of the reference layout
//    [e][6][nq^3], fetched by ONE cp.async.bulk (TMA, SASS UBLKCP) into shared memory, completion
//    on an mbarrier.  The copy for batch b+1 is issued as soon as the flux loop of batch b has
//    drained the buffer; 3-6 resident CTAs per SM keep > 100 KB in flight per SM with zero
//    registers spent on it.
//  * every contraction is "register column x constant-bank matrix": besides layout P (thread =
//    (q,r), column over p) the threads of an element also act in layout Q (thread = (p,r), column
//    over q) and layout R (thread = (p,q), column over r).  A layout change costs one shared-memory
//    store + one load per point instead of nq loads per point: 22 accesses per point instead of
//    4 nq + 9.  Two staging arrays with different paddings make every access conflict-free:
//      RQ[p][q][r] plane stride = nq^2 padded to == nq (mod 16)   (column reads with r across lanes)
//      RR[p][q][r] row stride odd                                  (row reads with q across lanes)
//  * results (flux, transposed derivative) are written back in place, so two arrays per element
//    suffice (three with interpolation).
//
// Numerics contract: <= 1e-12 relative max-norm against the oracle (tests/test_bk_gpu.py, tests/test_operator_gpu.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sumfact.cuh"

namespace b200fe {

namespace v2 {

constexpr __host__ __device__ int pad_plane_q(int nq)
{
#ifdef B200FE_V2_NOPAD_MAXNQ
    // tuning variant: tiny planes (p = 1, 2) trade the conflict-free padding (9 -> 19 doubles at nq = 3) for occupancy
    if (nq <= B200FE_V2_NOPAD_MAXNQ) return nq * nq;
#endif
    // smallest PS >= nq^2 with PS == nq (mod 16)
    int ps = nq * nq;
    while ((ps - nq) % 16 != 0) ++ps;
    return ps;
}

// ---- compile-time search of the [i][j][r] staging strides (interpolated operators) -------------
// Two access patterns hit that array: rows (thread = (i,j), slot = i*PA + j*RA) and columns
// (thread = (i,r), slot = i*PA + r).  A 64-bit shared-memory access is conflict-free when the 16
// threads of a half-warp fall into 16 different 8-byte slots; cost = sum over half-warps of the
// worst multiplicity.  (tools: /tmp-style python prototype reproduced these numbers.)
constexpr int halfwarp_cost(int n_threads, int n_active, int div, int s_hi, int s_lo)
{
    int total = 0;
    for (int h = 0; h < n_threads; h += 16) {
        int cnt[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        int worst = 0;
        for (int t = h; t < h + 16 && t < n_active; ++t) {
            const int slot = ((t / div) * s_hi + (t % div) * s_lo) % 16;
            if (++cnt[slot] > worst) worst = cnt[slot];
        }
        total += worst;
    }
    return total;
}
constexpr int stride_cost(int nm, int nq, int pa, int ra)
{
    return nq * halfwarp_cost(nq * nq, nm * nm, nm, pa, ra) + nm * halfwarp_cost(nq * nq, nm * nq, nq, pa, 1);
}
struct Strides { int ra, pa; };
constexpr Strides best_strides(int nm, int nq)
{
    Strides best{odd(nq), nm * odd(nq)};
    int best_cost = stride_cost(nm, nq, best.pa, best.ra);
    for (int ra = nq; ra <= nq + 4; ++ra)
        for (int pa = nm * ra; pa <= nm * ra + 16; ++pa) {
            const int c = stride_cost(nm, nq, pa, ra);
            if (c < best_cost || (c == best_cost && nm * pa < nm * best.pa)) { best_cost = c; best = Strides{ra, pa}; }
        }
    return best;
}

template <int NM, int NQ, bool COLL, int QOP>
struct Layout2 {
    static constexpr int N2 = NQ * NQ, N3 = N2 * NQ, M3 = NM * NM * NM;
    static constexpr bool LAP = (QOP & QOP_LAPLACE) != 0;
    static constexpr int PSQ = pad_plane_q(NQ);       // RQ plane stride
    static constexpr int RSR = odd(NQ);               // RR row stride
    static constexpr int PSR = NQ * RSR;              // RR plane stride
    static constexpr Strides SA = best_strides(NM, NQ);
    static constexpr int RA = SA.ra, PA = SA.pa;      // A / Y staging [i][j][r]: row and plane stride
    static constexpr int PB = PSQ;                    // B~ / X staging [i][q][r]: dense rows, padded planes
    static constexpr int SZ_FLUX = cmax(NQ * PSQ, NQ * PSR);
    static constexpr int SZ_INTERP = cmax(cmax(NM * PA, NM * PB), NM * NM * odd(NM));  // A|Y, B~|X, U|Z
    static constexpr int REGION = (COLL ? SZ_FLUX : cmax(SZ_FLUX, SZ_INTERP) + 1) & ~1;  // even: keeps 16 B alignment
    static constexpr int N_REGIONS = COLL ? (LAP ? 2 : 0) : 3;
    // element stride == nq^2 (mod 16): a warp that spans two elements keeps hitting distinct banks
    static constexpr int pad_elem(int w) { while ((w - N2) % 16 != 0) ++w; return w; }
    static constexpr int WORK_PER_ELEM = N_REGIONS == 0 ? 0 : pad_elem(N_REGIONS * REGION);
    static constexpr bool AFFINE = (QOP & QOP_AFFINE) != 0;  // no streamed G: nothing to stage
    static constexpr bool TRILIN = (QOP & QOP_TRILINEAR) != 0;
    static constexpr int G_PER_ELEM = (LAP && !AFFINE && !TRILIN) ? 6 * N3 : 0;
    // warp-local mode (nq <= 5): an element's nq^2 threads sit inside one warp, EPW elements per warp;
    // every intra-element barrier becomes __syncwarp().  Idle lanes of a warp work on a dummy slot.
    // Measured (profiles/r01d_warp_local.txt): a win only when the planes tile the warp exactly (nq = 2, 4:
    // bk3 p=2 72 -> 82 %, bk1 p=2 44 -> 58 %); with idle lanes (nq = 3, 5) the CTA-barrier version is faster.
    // Collocated kernels (4 barriers per batch) lose slightly (bk5 p=3 86 -> 82 %), so interpolated operators only.
    static constexpr bool WARP_LOCAL = !COLL && N2 <= 32 && 32 % N2 == 0 && WORK_PER_ELEM > 0;
    static constexpr int EPW = WARP_LOCAL ? 32 / N2 : 0;
    static constexpr size_t smem_bytes(int epb)
    {
        return 16 /* mbarrier */ + sizeof(double) * ((size_t)epb * G_PER_ELEM + (size_t)(epb + (WARP_LOCAL ? 1 : 0)) * WORK_PER_ELEM);
    }
    static constexpr int threads(int epb) { return WARP_LOCAL ? (epb / EPW) * 32 : epb * N2; }
};

// ---- PTX helpers (mbarrier + bulk async copy) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// out[m] = sum_n M(m,n) in[n] with M(m,n) = mat[m*RS + n*CS] (compile-time indices -> constant bank)
template <int NOUT, int NIN, int RS, int CS, typename MatT>
__device__ __forceinline__ void col_mul(const MatT &mat, const double (&in)[NIN], double (&out)[NOUT])
{
#pragma unroll
    for (int m = 0; m < NOUT; ++m) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < NIN; ++n) s = fma(mat[m * RS + n * CS], in[n], s);
        out[m] = s;
    }
}

// The four 1-D contractions of the kernel by name: plain register-column x constant-bank products, or (MatsT::kEvenOdd)
// the even-odd split of the symmetric 1-D matrices of a real basis (eo_contract.h: 41-46 % fewer DFMAs; measured on B200,
// profiles/r02a_*: BP3 p = 7 / 8 from 0.41 / 0.34 to 0.73 / 0.61 of the HBM roofline, spills gone).
template <int NM, int NQ, typename MatsT>
__device__ __forceinline__ void interp(const MatsT &m, const double (&in)[NM], double (&out)[NQ])
{
    if constexpr (MatsT::kEvenOdd) eo::interp<NM, NQ>(m.E, in, out);
    else col_mul<NQ, NM, NM, 1>(m.B, in, out);
}
template <int NM, int NQ, typename MatsT>
__device__ __forceinline__ void interp_t(const MatsT &m, const double (&in)[NQ], double (&out)[NM])
{
    if constexpr (MatsT::kEvenOdd) eo::interp_t<NM, NQ>(m.E, in, out);
    else col_mul<NM, NQ, 1, NM>(m.B, in, out);
}
template <int NM, int NQ, typename MatsT>
__device__ __forceinline__ void deriv(const MatsT &m, const double (&in)[NQ], double (&out)[NQ])
{
    if constexpr (MatsT::kEvenOdd) eo::deriv<NM, NQ>(m.E, in, out);
    else col_mul<NQ, NQ, NQ, 1>(m.D, in, out);
}
template <int NM, int NQ, typename MatsT>
__device__ __forceinline__ void deriv_t(const MatsT &m, const double (&in)[NQ], double (&out)[NQ])
{
    if constexpr (MatsT::kEvenOdd) eo::deriv_t<NM, NQ>(m.E, in, out);
    else col_mul<NQ, NQ, 1, NQ>(m.D, in, out);
}

}  // namespace v2

// ---------------------------------------------------------------------------------------------
// MC (multi-component): the batch loop gets an inner loop over the components of a component-blocked vector; the G block of
// a batch is waited for before the first and re-fetched after the last component, so the geometric factors (75-95 % of the
// bytes) are streamed once per cell instead of once per cell and component (BP6: 3 components).  Built for the collocated
// Laplace L-vector kernels; MC = false is the scalar kernel, unchanged.
template <int NM, int NQ, bool COLL, int QOP, bool LVEC, int EPB, int MINB, bool EO, bool MC = false>
__global__ void __launch_bounds__((v2::Layout2<NM, NQ, COLL, QOP>::threads(EPB)), MINB)
    sumfact2_kernel(const __grid_constant__ Mats<NM, NQ, EO> m, const KArgs a)
{
    static_assert(!MC || (COLL && LVEC && (QOP & QOP_LAPLACE) && !(QOP & QOP_MASS)), "multi-component kernels: collocated Laplace L-vector operators");
    using L = v2::Layout2<NM, NQ, COLL, QOP>;
    constexpr int N2 = L::N2, N3 = L::N3, M3 = L::M3, RA = L::RA, PA = L::PA, PB = L::PB;
    constexpr int PSQ = L::PSQ, RSR = L::RSR, PSR = L::PSR;
    constexpr bool LAP = L::LAP, MASS = (QOP & QOP_MASS) != 0, AFFINE = L::AFFINE, TRILIN = L::TRILIN, STORED_G = LAP && !AFFINE && !TRILIN;
    static_assert(!(AFFINE && TRILIN), "one on-the-fly geometry at a time");
    static_assert(!COLL || NM == NQ, "collocated operators need nm == nq");
    constexpr bool CART = (QOP & QOP_CARTESIAN) != 0;
    static_assert(!CART || (COLL && AFFINE && LVEC && !MASS && !MC), "cartesian kernels: collocated Laplace L-vector operators with per-cell constants");

    if (a.skip != nullptr && *a.skip != 0) return;  // uniform across the grid: decided before any barrier
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *Gs = reinterpret_cast<double *>(smem_raw + 16);      // [EPB][6][N3]
    double *work = Gs + EPB * L::G_PER_ELEM;
    const int tid = threadIdx.x;
    constexpr bool WL = L::WARP_LOCAL;
    // element slot and position in the plane; `lane_ok` is false only for the idle lanes of warp-local mode
    int el, t2;
    bool lane_ok = true;
    if constexpr (WL) {
        static_assert(!WL || EPB % (WL ? L::EPW : 1) == 0, "warp-local mode: EPB must be a multiple of EPW");
        const int warp = tid >> 5, lane = tid & 31;
        lane_ok = lane < L::EPW * N2;
        el = lane_ok ? warp * L::EPW + lane / N2 : EPB;  // EPB = dummy slot
        t2 = lane % N2;
    } else {
        el = tid / N2;
        t2 = tid - el * N2;
    }
    auto sync_elem = [&]() {  // barrier among the threads of one element
        if constexpr (WL) __syncwarp();
        else __syncthreads();
    };
    const int ta = t2 / NQ, tb = t2 - ta * NQ;  // (q,r) in layout P; (p,r) in layout Q; (p,q) in layout R
    double *R0 = work + el * L::WORK_PER_ELEM;
    [[maybe_unused]] double *R1 = R0 + L::REGION;
    [[maybe_unused]] double *R2 = R1 + L::REGION;
    double *RQ = R0, *RR = R1;
    const double *Ge = Gs + (lane_ok ? el : 0) * L::G_PER_ELEM + t2;

    const uint32_t n_batches = (a.n_elems + EPB - 1) / EPB;
    auto issue_g = [&](uint32_t eb) {  // one elected thread: fetch the batch's G block
        const uint32_t first = eb * EPB;
        const uint32_t cnt = (a.n_elems - first) < (uint32_t)EPB ? (a.n_elems - first) : (uint32_t)EPB;
        const uint32_t bytes = cnt * (uint32_t)(L::G_PER_ELEM * sizeof(double));
        v2::mbar_expect_tx(bar, bytes);
        v2::bulk_g2s(Gs, a.G + (size_t)first * L::G_PER_ELEM, bytes, bar);
    };
    if constexpr (STORED_G) {
        if (tid == 0) {
            v2::mbar_init(bar, 1);
            v2::fence_mbar_init();
        }
        // The unused slots of a tail batch read whatever the G buffer holds.  Zero it once, so that it only
        // ever contains zeros or earlier (finite) geometric factors: 0 * stale stays 0 in the fused inner
        // product without a branch in the flux loop (round-1 bug: 0 * NaN from a poisoned buffer).
        for (int i = tid; i < EPB * L::G_PER_ELEM; i += blockDim.x) Gs[i] = 0.0;
        v2::fence_proxy_async();  // order the generic-proxy zeros before the async-proxy (TMA) writes
        __syncthreads();
        if (tid == 0 && blockIdx.x < n_batches) issue_g(blockIdx.x);
    }
    uint32_t parity = 0;

    // Inputs are software-pipelined one batch ahead (ncu: the exposed in/idx -> src[idx] latency at the top
    // of a batch was the largest stall once G moved to TMA).  Collocated: a column over p per thread;
    // interpolated: a row over k for the threads (i,j) = t2 < NM^2, which feed the first sweep directly
    // from registers (no staging of U / Z in shared memory).
    // Interpolated operators: thread t2 of an element owns the nodal values l = t2 + c*nq^2 (c < NK), i.e. the
    // CTA reads / writes / gathers element vectors and index tables with fully coalesced requests; they are
    // re-shaped into rows through one shared-memory staging array.  (ncu, profiles/r01d_bk_l1tex.txt: per-thread
    // row access cost 19-20 sectors per request and made L1TEX the bottleneck of BK1/BK3 at 87-89 %.)
    // L-vector operators keep the register-row variant (thread (i,j) owns the row over k and feeds / drains the
    // outer sweeps without staging): measured faster there (BP3 p=5 0.65 vs 0.57, p=8 0.34 vs 0.29 of the HBM
    // roofline, profiles/r01f_*), while for element vectors both variants time the same.
    constexpr bool ROW_IO = LVEC && !COLL;
    constexpr int NK = (M3 + N2 - 1) / N2;
    constexpr int NIN = COLL ? NQ : (ROW_IO ? NM : NK);
    constexpr int RU = odd(NM);  // row stride of the nodal staging arrays U / Z
#ifndef B200FE_V2_PREFETCH_MAXNQ
#define B200FE_V2_PREFETCH_MAXNQ 10  // software-pipelined inputs for every degree (r01 sweep: needs >= 160 registers at nq >= 7)
#endif
    constexpr bool PREFETCH = NQ <= B200FE_V2_PREFETCH_MAXNQ;
    double cur_val[NIN];
    [[maybe_unused]] uint32_t cur_idx[NIN];
    [[maybe_unused]] double nxt_val[NIN];
    [[maybe_unused]] uint32_t nxt_idx[NIN];
    auto in_offset = [&](uint32_t e_, int n) -> size_t {
        if constexpr (ROW_IO) return (size_t)e_ * M3 + t2 * NM + n;   // row (i,j) = t2, entry k = n
        else return (size_t)e_ * (COLL ? N3 : M3) + n * N2 + t2;       // "plane n of the element, position t2"
    };
    auto in_valid = [&](int n) { return COLL || (ROW_IO ? t2 < NM * NM : n * N2 + t2 < M3); };
    auto load_idx = [&](uint32_t eb_, uint32_t (&ix)[NIN]) {
        const uint32_t e_ = eb_ * EPB + el;
        const bool ok = lane_ok && eb_ < n_batches && e_ < a.n_elems;
#pragma unroll
        for (int n = 0; n < NIN; ++n) ix[n] = (ok && in_valid(n)) ? __ldg(a.idx + in_offset(e_, n)) : kInvalidIndex;
    };
    auto load_val = [&](uint32_t eb_, const uint32_t (&ix)[NIN], double (&val)[NIN], const double *in_base = nullptr) {
        if constexpr (LVEC) {
            const double *src = MC ? in_base : a.in;
#pragma unroll
            for (int n = 0; n < NIN; ++n) val[n] = ix[n] == kInvalidIndex ? 0.0 : __ldg(src + ix[n]);
        } else {
            const uint32_t e_ = eb_ * EPB + el;
            const bool ok = lane_ok && eb_ < n_batches && e_ < a.n_elems;
#pragma unroll
            for (int n = 0; n < NIN; ++n) val[n] = (ok && in_valid(n)) ? __ldg(a.in + in_offset(e_, n)) : 0.0;
        }
    };
    if constexpr (LVEC) load_idx(blockIdx.x, cur_idx);
    load_val(blockIdx.x, cur_idx, cur_val, a.in);
    // KArgs::excl_interior: this thread's two fixed local indices are interior ones (the third runs over the register column)
    [[maybe_unused]] bool excl_t2 = false;
    if constexpr (LVEC) {
        const int t_hi = COLL ? ta : t2 / NM, t_lo = COLL ? tb : t2 % NM;
        excl_t2 = a.excl_interior != 0 && t_hi >= 1 && t_hi <= NM - 2 && t_lo >= 1 && t_lo <= NM - 2;
    }
    const int ncomp = MC ? a.ncomp : 1;
    constexpr bool JW_PIPE = MASS && !LAP && PREFETCH;
    [[maybe_unused]] double jw_cur[NQ];
    auto load_jw = [&](uint32_t eb_, double (&j)[NQ]) {
        const uint32_t e_ = eb_ * EPB + el;
        const bool ok = lane_ok && eb_ < n_batches && e_ < a.n_elems;
#pragma unroll
        for (int p = 0; p < NQ; ++p) j[p] = ok ? __ldg(a.JxW + (size_t)e_ * N3 + p * N2 + t2) : 0.0;
    };
    if constexpr (JW_PIPE) load_jw(blockIdx.x, jw_cur);

    double dot_acc = 0.0;
    for (uint32_t eb = blockIdx.x; eb < n_batches; eb += gridDim.x) {
        const uint32_t e = eb * EPB + el;
        const bool active = lane_ok && e < a.n_elems;
        const uint32_t nb = eb + gridDim.x;
      for (int comp = 0; comp < ncomp; ++comp) {  // (scalar kernels: one pass)
        const bool first_c = !MC || comp == 0, last_c = !MC || comp == ncomp - 1;
        [[maybe_unused]] const size_t coff = MC ? (size_t)comp * a.comp_stride : 0;
        // prefetch: next batch's indices (L-vector) or values (E-vector)
        if constexpr (PREFETCH) {
            if constexpr (LVEC) { if (first_c) load_idx(nb, nxt_idx); }
            else load_val(nb, nxt_idx, nxt_val);
        }
        [[maybe_unused]] double jw[NQ];
        if constexpr (MASS && JW_PIPE) {  // pure mass operator: JxW is the dominant stream -> one batch ahead, like `in`
#pragma unroll
            for (int p = 0; p < NQ; ++p) jw[p] = jw_cur[p];
            load_jw(nb, jw_cur);
        } else if constexpr (MASS) {  // Helmholtz: JxW column early, its latency hides behind the interpolation sweeps
            const double *Je = a.JxW + (size_t)(active ? e : 0) * N3 + t2;
#pragma unroll
            for (int p = 0; p < NQ; ++p) jw[p] = active ? __ldg(Je + p * N2) : 0.0;
        }
        double v[NQ];
        // ------------------------------------------------------------------ load (+ interpolation)
        if constexpr (COLL) {
#pragma unroll
            for (int p = 0; p < NQ; ++p) v[p] = cur_val[p];
        } else {
            if constexpr (ROW_IO) {
                if (t2 < NM * NM) {  // k -> r straight from the register row -> A[i][j][r] in R1
                    double u[NM], o[NQ];
#pragma unroll
                    for (int k = 0; k < NM; ++k) u[k] = cur_val[k];
                    v2::interp<NM, NQ>(m, u, o);
#pragma unroll
                    for (int r = 0; r < NQ; ++r) R1[(t2 / NM) * PA + (t2 % NM) * RA + r] = o[r];
                }
            } else {
#pragma unroll
                for (int c = 0; c < NK; ++c) {  // U -> R0, rows [i][j][.] with odd stride
                    const int l = t2 + c * N2;
                    if (l < M3) R0[(l / NM) * RU + (l % NM)] = cur_val[c];
                }
                sync_elem();
                if (t2 < NM * NM) {  // k -> r : rows of U -> A[i][j][r] in R1
                    double u[NM], o[NQ];
#pragma unroll
                    for (int k = 0; k < NM; ++k) u[k] = R0[t2 * RU + k];
                    v2::interp<NM, NQ>(m, u, o);
#pragma unroll
                    for (int r = 0; r < NQ; ++r) R1[(t2 / NM) * PA + (t2 % NM) * RA + r] = o[r];
                }
            }
            sync_elem();
            if (t2 < NM * NQ) {  // j -> q : columns of A -> B~[i][q][r] in R2
                const int i = t2 / NQ, r2 = t2 - i * NQ;
                double u[NM], o[NQ];
#pragma unroll
                for (int j = 0; j < NM; ++j) u[j] = R1[i * PA + j * RA + r2];
                v2::interp<NM, NQ>(m, u, o);
#pragma unroll
                for (int q = 0; q < NQ; ++q) R2[i * PB + q * NQ + r2] = o[q];
            }
            sync_elem();
            {  // i -> p into the register column
                double u[NM];
#pragma unroll
                for (int i = 0; i < NM; ++i) u[i] = R2[i * PB + t2];
                v2::interp<NM, NQ>(m, u, v);
            }
        }

        // ------------------------------------------------------------------ operator at the points
        double w[NQ];
        if constexpr (CART) {
            // axis-aligned cells (QOP_CARTESIAN): out = c_rr w_q w_r (S v)_p + w_p (c_ss w_r (S v)_q + c_tt w_q (S v)_r), S = m.B
            // (the three cell constants are fetched first: their latency hides behind the contractions)
            const double *c8 = a.cellG + (size_t)(active ? e : 0) * 8;
            const double c_rr = active ? __ldg(c8 + 0) : 0.0, c_ss = active ? __ldg(c8 + 3) : 0.0, c_tt = active ? __ldg(c8 + 5) : 0.0;
            double sp[NQ];
            v2::interp<NM, NQ>(m, v, sp);  // S along p from the register column
#pragma unroll
            for (int p = 0; p < NQ; ++p) {
                RQ[p * PSQ + t2] = v[p];
                RR[p * PSR + ta * RSR + tb] = v[p];
            }
            sync_elem();
            {   // layout Q: thread (p,r) = (ta,tb), column over q, in place
                double c[NQ], o[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) c[q] = RQ[ta * PSQ + q * NQ + tb];
                v2::interp<NM, NQ>(m, c, o);
#pragma unroll
                for (int q = 0; q < NQ; ++q) RQ[ta * PSQ + q * NQ + tb] = o[q];
            }
            {   // layout R: thread (p,q) = (ta,tb), row over r, in place
                double c[NQ], o[NQ];
#pragma unroll
                for (int r = 0; r < NQ; ++r) c[r] = RR[ta * PSR + tb * RSR + r];
                v2::interp<NM, NQ>(m, c, o);
#pragma unroll
                for (int r = 0; r < NQ; ++r) RR[ta * PSR + tb * RSR + r] = o[r];
            }
            sync_elem();
            const double wq = m.W[ta], wr = m.W[tb];
            const double kr = c_rr * wq * wr, ks = c_ss * wr, kt = c_tt * wq;
            if constexpr (PREFETCH) load_val(nb, nxt_idx, nxt_val, a.in);  // next batch's gathers (indices arrived long ago)
#pragma unroll
            for (int p = 0; p < NQ; ++p) {
                w[p] = fma(kr, sp[p], m.W[p] * fma(ks, RQ[p * PSQ + t2], kt * RR[p * PSR + ta * RSR + tb]));
                dot_acc = fma(v[p], w[p], dot_acc);  // u.(A u): collocated, so the nodal values are the point values
            }
        } else if constexpr (LAP) {
            [[maybe_unused]] double cg[6];
            if constexpr (AFFINE) {  // six constants per cell (scaled by the tensor-product quadrature weight in the flux loop);
                                     // fetched first: the latency hides behind the derivative sweeps
                const double *c8 = a.cellG + (size_t)(active ? e : 0) * 8;
#pragma unroll
                for (int c = 0; c < 6; ++c) cg[c] = active ? __ldg(c8 + c) : 0.0;
            }
            double gr[NQ];  // d/dr (along p) from the register column; later the r-flux
            v2::deriv<NM, NQ>(m, v, gr);
#pragma unroll
            for (int p = 0; p < NQ; ++p) {
                RQ[p * PSQ + t2] = v[p];
                RR[p * PSR + ta * RSR + tb] = v[p];
            }
            sync_elem();
            {   // layout Q: thread (p,r) = (ta,tb), column over q, in place
                double c[NQ], o[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) c[q] = RQ[ta * PSQ + q * NQ + tb];
                v2::deriv<NM, NQ>(m, c, o);
#pragma unroll
                for (int q = 0; q < NQ; ++q) RQ[ta * PSQ + q * NQ + tb] = o[q];
            }
            {   // layout R: thread (p,q) = (ta,tb), row over r, in place
                double c[NQ], o[NQ];
#pragma unroll
                for (int r = 0; r < NQ; ++r) c[r] = RR[ta * PSR + tb * RSR + r];
                v2::deriv<NM, NQ>(m, c, o);
#pragma unroll
                for (int r = 0; r < NQ; ++r) RR[ta * PSR + tb * RSR + r] = o[r];
            }
            sync_elem();
            if constexpr (STORED_G) {
                if (first_c) {
                    v2::mbar_wait(bar, parity);  // G of this batch has landed (all components use it)
                    parity ^= 1u;
                }
            }
            [[maybe_unused]] double wqr = 0.0;
            // trilinear cells: the columns of J at this thread's (x^, y^) as linear functions of z^ (J[.][2] is constant in z^)
            [[maybe_unused]] double j0a[3], j0b[3], j1a[3], j1b[3], j2c[3];
            if constexpr (TRILIN) {
                const double *X = a.cellG + (size_t)(active ? e : 0) * 24;
                const double xx = m.X[tb], xy = m.X[ta];  // x^ = fastest local index (tb), y^ = middle (ta)
                wqr = m.W[ta] * m.W[tb];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double v8[8];  // vertex index c*4 + b*2 + a  (a <-> x^)
#pragma unroll
                    for (int k = 0; k < 8; ++k) v8[k] = active ? __ldg(X + d * 8 + k) : (((k >> d) & 1) ? 1.0 : 0.0);  // idle slots: unit cube
                    // d/dx^: differences along a, bilinear in (y^, z^)
                    const double dx0 = fma(xy, (v8[3] - v8[2]) - (v8[1] - v8[0]), v8[1] - v8[0]);   // z^ = 0
                    const double dx1 = fma(xy, (v8[7] - v8[6]) - (v8[5] - v8[4]), v8[5] - v8[4]);   // z^ = 1
                    j0a[d] = dx0; j0b[d] = dx1 - dx0;
                    // d/dy^: differences along b, bilinear in (x^, z^)
                    const double dy0 = fma(xx, (v8[3] - v8[1]) - (v8[2] - v8[0]), v8[2] - v8[0]);
                    const double dy1 = fma(xx, (v8[7] - v8[5]) - (v8[6] - v8[4]), v8[6] - v8[4]);
                    j1a[d] = dy0; j1b[d] = dy1 - dy0;
                    // d/dz^: differences along c, bilinear in (x^, y^): constant along the column
                    const double dz0 = fma(xx, (v8[5] - v8[1]) - (v8[4] - v8[0]), v8[4] - v8[0]);   // y^ = 0
                    const double dz1 = fma(xx, (v8[7] - v8[3]) - (v8[6] - v8[2]), v8[6] - v8[2]);   // y^ = 1
                    j2c[d] = fma(xy, dz1 - dz0, dz0);
                }
            }
            if constexpr (AFFINE) wqr = m.W[ta] * m.W[tb];
#pragma unroll
            for (int p = 0; p < NQ; ++p) {
                const double qr = gr[p];
                const double qs = RQ[p * PSQ + t2];
                const double qt = RR[p * PSR + ta * RSR + tb];
                double g0, g1, g2, g3, g4, g5;
                if constexpr (TRILIN) {
                    // J[d][b] = d x_d / d xi_b at (x^, y^, z^ = X[p]); adj = det * J^-1; G = w / det * adj adj^T with the kernel's
                    // directions (r,s,t) = (z^, y^, x^)  (geometry_kernel of operator.cu, the same arithmetic per point)
                    const double xz = m.X[p];
                    double J[3][3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        J[d][0] = fma(xz, j0b[d], j0a[d]);
                        J[d][1] = fma(xz, j1b[d], j1a[d]);
                        J[d][2] = j2c[d];
                    }
                    // rows of adj: A[b][a] = cofactor, so that K[b][a] = A[b][a] / det = d xi_b / d x_a
                    double A[3][3];
                    A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
                    A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
                    A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
                    A[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
                    A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
                    A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
                    A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
                    A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
                    A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
                    const double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
                    const double sc = wqr * m.W[p] / det;
                    // (r,s,t) = rows 2,1,0 of K
                    g0 = sc * (A[2][0] * A[2][0] + A[2][1] * A[2][1] + A[2][2] * A[2][2]);
                    g1 = sc * (A[2][0] * A[1][0] + A[2][1] * A[1][1] + A[2][2] * A[1][2]);
                    g2 = sc * (A[2][0] * A[0][0] + A[2][1] * A[0][1] + A[2][2] * A[0][2]);
                    g3 = sc * (A[1][0] * A[1][0] + A[1][1] * A[1][1] + A[1][2] * A[1][2]);
                    g4 = sc * (A[1][0] * A[0][0] + A[1][1] * A[0][1] + A[1][2] * A[0][2]);
                    g5 = sc * (A[0][0] * A[0][0] + A[0][1] * A[0][1] + A[0][2] * A[0][2]);
                } else if constexpr (AFFINE) {
                    const double wt = wqr * m.W[p];
                    g0 = cg[0] * wt; g1 = cg[1] * wt; g2 = cg[2] * wt; g3 = cg[3] * wt; g4 = cg[4] * wt; g5 = cg[5] * wt;
                } else {
                    g0 = Ge[0 * N3 + p * N2]; g1 = Ge[1 * N3 + p * N2]; g2 = Ge[2 * N3 + p * N2];
                    g3 = Ge[3 * N3 + p * N2]; g4 = Ge[4 * N3 + p * N2]; g5 = Ge[5 * N3 + p * N2];
                }
                const double fr = g0 * qr + g1 * qs + g2 * qt;
                const double fs = g1 * qr + g3 * qs + g4 * qt;
                const double ft = g2 * qr + g4 * qs + g5 * qt;
                gr[p] = fr;
                RQ[p * PSQ + t2] = fs;
                RR[p * PSR + ta * RSR + tb] = ft;
                // u.(A u) = sum over points of grad(u)^T G grad(u): the CG inner product comes for free here,
                // no need to keep the gathered values alive until the scatter
                // (inactive slots contribute 0 * finite: the G buffer is zero-initialised, see above)
                if constexpr (LVEC) dot_acc = fma(qr, fr, fma(qs, fs, fma(qt, ft, dot_acc)));
            }
            if constexpr (STORED_G) __syncthreads();  // CTA-wide: the shared G buffer is about to be refilled
            else sync_elem();
            if constexpr (STORED_G) {
                if (tid == 0 && last_c) {  // the G buffer is drained: fetch the next batch's block behind the rest of this one
                    if (nb < n_batches) {
                        v2::fence_proxy_async();
                        issue_g(nb);
                    }
                }
            }
            if constexpr (LVEC && PREFETCH) {  // next item's gathers (indices arrived long ago): next component, or next batch
                if constexpr (MC) {
                    if (last_c) load_val(nb, nxt_idx, nxt_val, a.in);
                    else load_val(eb, cur_idx, nxt_val, a.in + coff + a.comp_stride);
                } else load_val(nb, nxt_idx, nxt_val, a.in);
            }
            v2::deriv_t<NM, NQ>(m, gr, w);  // w[p'] = sum_p D[p][p'] f_r[p]
            {   // layout Q, transposed derivative along q, in place
                double c[NQ], o[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) c[q] = RQ[ta * PSQ + q * NQ + tb];
                v2::deriv_t<NM, NQ>(m, c, o);
#pragma unroll
                for (int q = 0; q < NQ; ++q) RQ[ta * PSQ + q * NQ + tb] = o[q];
            }
            {   // layout R, transposed derivative along r, in place
                double c[NQ], o[NQ];
#pragma unroll
                for (int r = 0; r < NQ; ++r) c[r] = RR[ta * PSR + tb * RSR + r];
                v2::deriv_t<NM, NQ>(m, c, o);
#pragma unroll
                for (int r = 0; r < NQ; ++r) RR[ta * PSR + tb * RSR + r] = o[r];
            }
            sync_elem();
#pragma unroll
            for (int p = 0; p < NQ; ++p) w[p] += RQ[p * PSQ + t2] + RR[p * PSR + ta * RSR + tb];
        } else {
#pragma unroll
            for (int p = 0; p < NQ; ++p) w[p] = 0.0;
        }
        if constexpr (MASS) {
#pragma unroll
            for (int p = 0; p < NQ; ++p) {
                const double mv = jw[p] * v[p];
                w[p] += mv;
                if constexpr (LVEC) dot_acc = fma(mv, v[p], dot_acc);  // + u^T M u at the points (jw = 0 on inactive slots)
            }
        }
        if constexpr (LVEC && !LAP && PREFETCH) load_val(nb, nxt_idx, nxt_val, a.in);

        // ------------------------------------------------------------------ (interpolation +) store
        if constexpr (COLL) {
            if constexpr (LVEC) {
#pragma unroll
                for (int p = 0; p < NQ; ++p) {
                    if (cur_idx[p] != kInvalidIndex) {
                        if (excl_t2 && p >= 1 && p <= NQ - 2) a.out[coff + cur_idx[p]] = w[p];  // cell-interior DoF: sole writer
                        else atomicAdd(a.out + coff + cur_idx[p], w[p]);
                    }
                }
            } else {
                if (active) {
#pragma unroll
                    for (int p = 0; p < NQ; ++p) a.out[(size_t)e * N3 + p * N2 + t2] = w[p];
                }
            }
        } else {
            {   // p -> i in registers; X[i][q][r] -> R2 (dense)
                double x[NM];
                v2::interp_t<NM, NQ>(m, w, x);
#pragma unroll
                for (int i = 0; i < NM; ++i) R2[i * PB + t2] = x[i];
            }
            sync_elem();
            if (t2 < NM * NQ) {  // q -> j : columns of X -> Y[i][j][r] in R0
                const int i = t2 / NQ, r2 = t2 - i * NQ;
                double x[NQ], o[NM];
#pragma unroll
                for (int q = 0; q < NQ; ++q) x[q] = R2[i * PB + q * NQ + r2];
                v2::interp_t<NM, NQ>(m, x, o);
#pragma unroll
                for (int j = 0; j < NM; ++j) R0[i * PA + j * RA + r2] = o[j];
            }
            sync_elem();
            if (t2 < NM * NM) {  // r -> k : rows of Y -> Z[i][j][k]
                double x[NQ], z[NM];
#pragma unroll
                for (int r = 0; r < NQ; ++r) x[r] = R0[(t2 / NM) * PA + (t2 % NM) * RA + r];
                v2::interp_t<NM, NQ>(m, x, z);
                if constexpr (ROW_IO) {  // scatter the row straight from registers
#pragma unroll
                    for (int k = 0; k < NM; ++k)
                        if (cur_idx[k] != kInvalidIndex) {
                            if (excl_t2 && k >= 1 && k <= NM - 2) a.out[cur_idx[k]] = z[k];  // cell-interior DoF: sole writer
                            else atomicAdd(a.out + cur_idx[k], z[k]);
                        }
                } else {
#pragma unroll
                    for (int k = 0; k < NM; ++k) R1[t2 * RU + k] = z[k];  // odd row stride
                }
            }
            if constexpr (!ROW_IO) sync_elem();
#pragma unroll
            for (int c = 0; c < (ROW_IO ? 0 : NK); ++c) {  // coalesced store of the nodal values this thread owns
                const int l = t2 + c * N2;
                if (l < M3) {
                    const double z = R1[(l / NM) * RU + (l % NM)];
                    if constexpr (LVEC) {
                        if (cur_idx[c] != kInvalidIndex) atomicAdd(a.out + cur_idx[c], z);
                    } else {
                        if (active) a.out[(size_t)e * M3 + l] = z;
                    }
                }
            }
            // next batch: U -> R0 (last read by the r -> k sweep, fenced by the barrier above); its k-sweep
            // writes R1 only after the barrier that follows the U staging, i.e. after these Z reads
        }
        // rotate the software pipeline (or load the next batch's inputs now)
        if constexpr (PREFETCH) {
#pragma unroll
            for (int n = 0; n < NIN; ++n) {
                cur_val[n] = nxt_val[n];
                if constexpr (LVEC) { if (last_c) cur_idx[n] = nxt_idx[n]; }
            }
        } else {
            static_assert(!MC || PREFETCH, "multi-component kernels use the software-pipelined inputs");
            if constexpr (LVEC) load_idx(nb, cur_idx);
            load_val(nb, cur_idx, cur_val, a.in);
        }
      }  // components
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
