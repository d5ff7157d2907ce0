// oracle/ref_gpu_sumfact_mma.cu -- BENCH / TEST INFRASTRUCTURE ONLY (never linked into the product).
// extern "C" shim around the reference's FP64 tensor-core (DMMA, PTX mma.sync.m8n8k4.f64) BK1 kernel and its CUDA-core twin
// of the same study, compiled IN PLACE from /root/reference/sum_factorization/include for sm_100a:
//   BK1::Parallel::BwdTransHexKernel_mma<T, 8, 8, 4, nq, nq, nq>   sum_factorization/include/kernels/BK1/templated_cuda_mma_kernels.cuh:123-232
//     launch: one warp per CTA, numBlocks = (nelmt * 32 / 4) / 32   (sum_factorization/src/BK1/templated_cuda_mma_benchmark.cc:100-116,127-129)
//   BK1::Parallel::BwdTransHexKernel_QP_1D_Warp<T, nq, nq, nq>     .../BK1/templated_cuda_kernels.cuh:12-175 (one warp per element, CUDA cores)
// Output: oracle/_ref/libref_gpu_sumfact.so.  This is SURVEY's kernel K9: the evidence for "keep DMMA only if ncu shows it
// beating CUDA-core FMA".
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <limits>
#include <cuda_runtime.h>
#include <kernels/BK1/templated_cuda_kernels.cuh>
#include <kernels/BK1/templated_cuda_mma_kernels.cuh>

namespace {
using T = double;
template <unsigned nq>
cudaError_t launch(int variant, unsigned nelmt, unsigned numThreads, const T *basis, const T *JxW, const T *in, T *out, unsigned *shape)
{
    constexpr unsigned nm = nq - 1;
    if (variant == 0) {  // templated_cuda_mma_benchmark.cc:100-116: one warp per CTA
        unsigned numBlocks = numThreads / 32u;
        if (numBlocks == 0) numBlocks = 1;
        const size_t smem = (2 * nq * nq * nq + 3 * nm * nq) * sizeof(T);
        if (shape) { shape[0] = 1; shape[1] = numBlocks; shape[2] = 32; }
        BK1::Parallel::BwdTransHexKernel_mma<T, 8, 8, 4, nq, nq, nq><<<numBlocks, 32, smem>>>(nelmt, basis, basis, basis, JxW, in, out);
    } else {  // templated_cuda_benchmark.cc:99-124: as many warps (= elements) per CTA as 48 KB of shared memory / 512 threads allow
        int nelmtPerBlock = (int)((48 * 1024 / sizeof(T) - 3 * nq * nm) / (2 * nq * nq * nq));
        nelmtPerBlock = std::min(nelmtPerBlock, 512 / 32);
        unsigned grid = numThreads / (32u * nelmtPerBlock);
        if (grid == 0) grid = 1;
        const size_t smem = ((size_t)nelmtPerBlock * 2 * nq * nq * nq + 3 * nm * nq) * sizeof(T);
        if (shape) { shape[0] = (unsigned)nelmtPerBlock; shape[1] = grid; shape[2] = 32u * nelmtPerBlock; }
        BK1::Parallel::BwdTransHexKernel_QP_1D_Warp<T, nq, nq, nq><<<grid, 32 * nelmtPerBlock, smem>>>(nelmt, basis, basis, basis, JxW, in, out);
    }
    return cudaGetLastError();
}
}  // namespace

extern "C" {
// variant 0 = DMMA kernel, 1 = CUDA-core warp-per-element kernel.  Device pointers; JxW in the kernel's own layout
// ([e][r][q][p], p fastest -- templated_cuda_mma_kernels.cuh:196).  nq = 4 is the reference's only instantiation.
int ref_gpu_sumfact_bk1(int variant, int nq, unsigned nelmt, const double *d_basis, const double *d_JxW, const double *d_in,
                        double *d_out, int ntests, float *ms_min, float *ms_mean, unsigned *shape_out)
{
    if (nq != 4 && nq != 8) return (int)cudaErrorInvalidValue;
    const unsigned numThreads = nelmt * 32u / 4u;  // the drivers' default (templated_cuda_mma_benchmark.cc:128)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = std::numeric_limits<float>::max(), sum = 0.f;
    cudaError_t err = cudaSuccess;
    for (int t = 0; t < ntests + 2 && err == cudaSuccess; ++t) {
        cudaEventRecord(e0);
        err = nq == 4 ? launch<4>(variant, nelmt, numThreads, d_basis, d_JxW, d_in, d_out, shape_out) : launch<8>(variant, nelmt, numThreads, d_basis, d_JxW, d_in, d_out, shape_out);
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
