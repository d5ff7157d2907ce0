// common.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

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

}  // namespace b200fe
