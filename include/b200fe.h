/*
 * b200fe.h -- C ABI of the B200-native matrix-free operator path.
 *
 * Drop-in boundary for the hot path of dealii-X/benchmarks (paths below are relative to the
 * reference tree).  Every entry point takes plain pointers and sizes, returns an int status
 * (0 = success) and never throws.  Device pointers are named d_*, host pointers h_*.  Every
 * call that launches work takes the CUDA stream (cudaStream_t passed as void*) and performs no
 * hidden synchronisation unless its comment says so.  There is no CPU fallback: without a
 * CUDA device every compute call returns B200FE_ERR_CUDA.
 *
 * Index / layout conventions are the reference's own:
 *   element-local DoF  l = i*nm^2 + j*nm + k, k fastest (= x in deal.II lexicographic order)
 *   quadrature point   p*nq^2 + q*nq + r      (p <-> i, the slowest index)
 *   G  [cell][6][nq^3] components rr,rs,rt,ss,st,tt, r = direction of the slowest index
 *        (CEED_BK/include/kernels/BK3/templated_cuda_kernels.cuh:156-161,
 *         CEED_bp/include/portable_laplace_operator.h:294-295)
 *   JxW [cell][nq^3]
 */
#ifndef B200FE_H
#define B200FE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FE_VERSION 100 /* 0.1.0 */

enum {
    B200FE_OK = 0,
    B200FE_ERR_INVALID_ARG = 1, /* null pointer, size mismatch, bad enum */
    B200FE_ERR_UNSUPPORTED = 2, /* degree outside 1..8 or nq not in {p+1, p+2} */
    B200FE_ERR_CUDA = 3,        /* CUDA runtime error (text in b200fe_last_error) */
    B200FE_ERR_COMM = 4,        /* NCCL / peer-access error */
    B200FE_ERR_NO_CONVERGENCE = 5 /* CG hit max_it (deal.II SolverControl::NoConvergence) */
};

#define B200FE_INVALID_INDEX 0xFFFFFFFFu /* numbers::invalid_unsigned_int */

int b200fe_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char *b200fe_last_error(void);

/* ------------------------------------------------------------------------------------------
 * 1. Element-vector bake-off kernels (replace the <<<>>> launches of the standalone drivers).
 *    h_basis  [nq*nm]  basis[q*nm+i]           (CEED_BK/src/BK1/templated_cuda_benchmark.cc:52-58)
 *    h_dbasis [nq*nq]  dbasis[p*nq+n]          (CEED_BK/src/BK3/templated_cuda_benchmark.cc:60-66)
 *    The two small 1-D matrices are HOST pointers: they ride to the GPU as kernel parameters
 *    (constant bank), which replaces the reference's d_basis/d_dbasis device copies.
 * ------------------------------------------------------------------------------------------ */

/* out_e = B^T (JxW .* (B in_e)).  Replaces BK1::Parallel::MassOperator<T,nq><<<>>>
 * (CEED_BK/include/kernels/BK1/templated_cuda_kernels.cuh:10-195; launch at
 *  CEED_BK/src/BK1/templated_cuda_benchmark.cc:84).  nq may be p+1 or p+2. */
int b200fe_bk1_apply(int p, int nq, uint32_t nelmt, const double *h_basis, const double *d_JxW,
                     const double *d_in, double *d_out, void *stream);

/* out_e = B^T D^T G D B in_e.  Replaces BK3::Parallel::LaplaceOperator<T,nq><<<>>>
 * (CEED_BK/include/kernels/BK3/templated_cuda_kernels.cuh:11-292; launch at
 *  CEED_BK/src/BK3/templated_cuda_benchmark.cc:95). */
int b200fe_bk3_apply(int p, int nq, uint32_t nelmt, const double *h_basis, const double *h_dbasis,
                     const double *d_G, const double *d_in, double *d_out, void *stream);

/* out_e = D^T G D in_e, nm = nq = p+1.  Replaces BK5::Parallel::LaplaceOperator<T,nq><<<>>>
 * (CEED_BK/include/kernels/BK5/templated_cuda_kernels.cuh:11-126). */
int b200fe_bk5_apply(int p, uint32_t nelmt, const double *h_dbasis, const double *d_G,
                     const double *d_in, double *d_out, void *stream);

/* sum(x[i]^2) -> *d_result (device scalar, overwritten).  Replaces the thrust::transform_reduce
 * checksum of the drivers (CEED_BK/src/BK3/templated_cuda_benchmark.cc:104-108). */
int b200fe_sum_squares(uint64_t n, const double *d_x, double *d_result, void *stream);

/* Launch configuration actually used by the E-/L-vector kernels (for the benchmark tables that
 * print nelmtPerBatch / numBlocks / threadsPerBlock, CEED_BK/include/benchmark_printer.hpp:8-81).
 * kind: 1 = BK1, 3 = BK3, 5 = BK5. */
int b200fe_bk_launch_info(int kind, int p, int nq, uint32_t nelmt, int *elems_per_block,
                          int *num_blocks, int *threads_per_block, int *smem_bytes);

#ifdef __cplusplus
}
#endif
#endif /* B200FE_H */
