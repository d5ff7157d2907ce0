// oracle/ref_gpu_ceedbk.cu -- BENCH / TEST INFRASTRUCTURE ONLY (never linked into the product).
// extern "C" shim around the reference's OWN raw-CUDA kernels ("the kernel to beat", SURVEY 2b / BASELINE.md 2.4),
// compiled IN PLACE from /root/reference/CEED_BK/include for sm_100a with T = double (the reference driver's
// `using T = float`, CEED_BK/src/BK3/templated_cuda_benchmark.cc:125, switched to the FP64 the metric is quoted in).
// Output: oracle/_ref/libref_gpu_ceedbk.so (git-ignored, travels to the GPU box).
//   BK1::Parallel::MassOperator     CEED_BK/include/kernels/BK1/templated_cuda_kernels.cuh:10-195
//   BK3::Parallel::LaplaceOperator  CEED_BK/include/kernels/BK3/templated_cuda_kernels.cuh:11-292
//   BK5::Parallel::LaplaceOperator  CEED_BK/include/kernels/BK5/templated_cuda_kernels.cuh:11-126
// Launch shape = the reference drivers' defaults (CEED_BK/src/BK{1,3,5}/templated_cuda_benchmark.cc, main()):
//   shmemPerBlock = 10800; nelmtPerBatch = shmemPerBlock / (c nq^3) / sizeof(T) (c = 2 for BK1, 4 for BK3/BK5), at least 1;
//   numBlocks = ceil(nelmt / nelmtPerBatch) / 2; threadsPerBlock = nq^2 * nelmtPerBatch; dynamic shared memory as in run_test().
// Timing: CUDA events around each launch (the reference times launch + cudaDeviceSynchronize with a host clock and keeps
// the minimum, :93-104); min and mean over ntests are returned.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <limits>
#include <cuda_runtime.h>
#include <kernels/BK1/templated_cuda_kernels.cuh>
#include <kernels/BK3/templated_cuda_kernels.cuh>
#include <kernels/BK5/templated_cuda_kernels.cuh>

namespace {
using T = double;
struct Shape { unsigned nelmtPerBatch, numBlocks, threads; size_t smem; };

Shape shape_for(int kind, unsigned nq, unsigned nelmt)
{
    const unsigned nm = nq - 1;
    const int shmemPerBlock = 10800;
    const unsigned c = kind == 1 ? 2u : 4u;
    unsigned epb = shmemPerBlock / (c * nq * nq * nq) / sizeof(T);
    if (epb == 0) epb = 1;
    unsigned nb = (nelmt + epb - 1) / epb / 2;
    if (nb == 0) nb = 1;
    const unsigned threads = nq * nq * std::max(1u, epb);
    size_t ssize = 0;
    if (kind == 1) ssize = nm * nq + 2 * epb * nq * nq * nq;
    if (kind == 3) ssize = nm * nq + nq * nq + 4 * epb * nq * nq * nq;
    if (kind == 5) ssize = nq * nq + 4 * epb * nq * nq * nq;
    return Shape{epb, nb, threads, ssize * sizeof(T)};
}

template <unsigned nq>
cudaError_t launch(int kind, unsigned nelmt, const Shape &s, const T *basis, const T *dbasis, const T *geom, const T *in, T *out)
{
    if (kind == 1) {
        if constexpr (nq >= 3) BK1::Parallel::MassOperator<T, nq><<<s.numBlocks, s.threads, s.smem>>>(nelmt, s.nelmtPerBatch, basis, geom, in, out);
    } else if (kind == 3) {
        if constexpr (nq >= 3) BK3::Parallel::LaplaceOperator<T, nq><<<s.numBlocks, s.threads, s.smem>>>(nelmt, s.nelmtPerBatch, basis, dbasis, geom, in, out);
    } else {
        if constexpr (nq <= 9) BK5::Parallel::LaplaceOperator<T, nq><<<s.numBlocks, s.threads, s.smem>>>(nelmt, s.nelmtPerBatch, dbasis, geom, in, out);
    }
    return cudaGetLastError();
}

cudaError_t dispatch(int kind, unsigned nq, unsigned nelmt, const Shape &s, const T *basis, const T *dbasis, const T *geom, const T *in, T *out)
{
    switch (nq) {
    case 2: return launch<2>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 3: return launch<3>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 4: return launch<4>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 5: return launch<5>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 6: return launch<6>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 7: return launch<7>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 8: return launch<8>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 9: return launch<9>(kind, nelmt, s, basis, dbasis, geom, in, out);
    case 10: return launch<10>(kind, nelmt, s, basis, dbasis, geom, in, out);
    }
    return cudaErrorInvalidValue;
}
}  // namespace

extern "C" {
// kind = 1 | 3 | 5; nq = p + 2 (BK1, BK3) or p + 1 (BK5).  All pointers are DEVICE pointers: basis [nq*nm], dbasis [nq*nq],
// geom = JxW [nelmt*nq^3] (BK1) or G [nelmt*6*nq^3] (BK3/BK5), in / out element vectors.  shape_out[3] = nelmtPerBatch,
// numBlocks, threadsPerBlock.  Returns a cudaError_t value (0 = success).
int ref_gpu_ceedbk(int kind, int nq, unsigned nelmt, const double *d_basis, const double *d_dbasis, const double *d_geom,
                   const double *d_in, double *d_out, int ntests, float *ms_min, float *ms_mean, unsigned *shape_out)
{
    if ((kind != 1 && kind != 3 && kind != 5) || nq < 2 || nq > 10 || (kind != 5 && nq < 3) || (kind == 5 && nq > 9)) return (int)cudaErrorInvalidValue;
    const Shape s = shape_for(kind, (unsigned)nq, nelmt);
    if (shape_out) { shape_out[0] = s.nelmtPerBatch; shape_out[1] = s.numBlocks; shape_out[2] = s.threads; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = std::numeric_limits<float>::max(), sum = 0.f;
    cudaError_t err = cudaSuccess;
    for (int t = 0; t < ntests + 2 && err == cudaSuccess; ++t) {  // two untimed warm-up launches
        cudaEventRecord(e0);
        err = dispatch(kind, (unsigned)nq, nelmt, s, d_basis, d_dbasis, d_geom, d_in, d_out);
        cudaEventRecord(e1);
        if (err == cudaSuccess) err = cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (t >= 2) { best = std::min(best, ms); sum += ms; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_min) *ms_min = best;
    if (ms_mean) *ms_mean = ntests > 0 ? sum / ntests : 0.f;
    return (int)err;
}
}
