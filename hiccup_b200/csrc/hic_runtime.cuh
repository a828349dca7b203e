// hic_runtime.cuh -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/hiccup_b200.h"

namespace hic {

char* last_error_buffer();            // thread-local, defined in hic_runtime.cu
int fail(int code, const char* fmt, ...);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// optional per-kernel timing with CUDA events on the launching stream (hic_profile_* in the ABI)
void prof_begin(const char* name, cudaStream_t st);
void prof_end(cudaStream_t st);

}  // namespace hic

#define HIC_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return hic::fail(HIC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    } while (0)

#define HIC_CHECK_LAUNCH(name)                                                                \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return hic::fail(HIC_ERR_CUDA, "launch of %s failed: %s", name,                   \
                             cudaGetErrorString(_e));                                         \
    } while (0)

#define HIC_LAUNCH(name, st, ...)                                                             \
    do {                                                                                      \
        hic::prof_begin(name, st);                                                            \
        __VA_ARGS__;                                                                          \
        hic::prof_end(st);                                                                    \
        HIC_CHECK_LAUNCH(name);                                                               \
    } while (0)

#define HIC_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) return hic::fail(HIC_ERR_INVALID, __VA_ARGS__);                          \
    } while (0)
