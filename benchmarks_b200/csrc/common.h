// common.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX 3: no-ops unless a profiler (nsys / ncu --nvtx) injects itself

#include <cstdarg>
#include <cstdio>

#include "../../include/b200fe.h"

namespace b200fe {

// thread-local last-error text (returned by b200fe_last_error)
char *error_buffer();
int fail(int code, const char *fmt, ...);
int fail_cuda(cudaError_t e, const char *what);

#define B200FE_CUDA_TRY(expr)                                            \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return ::b200fe::fail_cuda(_e, #expr);    \
    } while (0)

#define B200FE_REQUIRE(cond, ...)                                                        \
    do {                                                                                 \
        if (!(cond)) return ::b200fe::fail(B200FE_ERR_INVALID_ARG, __VA_ARGS__);        \
    } while (0)

// Named host-side range for nsys / ncu timelines.  The names follow the reference's own markers: LIKWID regions
// "cg_solver" / "matvec" (bp5_kokkos/benchmark.cc:358-398) and the Kokkos kernel labels of the deal.II cell loop
// (bakeoff_problems_dealii/include/portable_laplace_operator.h:663).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

}  // namespace b200fe
