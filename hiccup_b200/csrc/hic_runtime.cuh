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

// Small transfers through the SMs.  A copy engine serves its queue strictly in order (tools/ce_fifo.py: a
// 64 KB copy waits the full ~2 ms behind a 105 MB image upload that another CUDA stream queued first), so
// in the pipelined batch path every code table, offset array or status word would stall behind whole
// image batches.  These helpers move small data with a copy kernel that reads or writes page-locked
// host memory directly (zero-copy over PCIe): no copy engine, no queueing behind bulk transfers.
// Pageable sources are first copied into the plan's page-locked staging area.
struct SmallXfer {
    char* h_stage = nullptr;        // page-locked staging, reused by every call of the owning plan
    size_t cap = 0, used = 0;
    void reset() { used = 0; }
    void destroy();
};
constexpr size_t SMALL_XFER_MAX = 8u << 20;
// host -> device on `st`; the host data may go as soon as the call returns only if it was staged (pageable
// source) -- a page-locked source is read when the kernel runs, like cudaMemcpyAsync
int small_h2d(SmallXfer& x, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
// device -> page-locked staging on `st`; *h_where is valid after the stream has been synchronised
int small_d2h(SmallXfer& x, const void* d_src, size_t bytes, cudaStream_t st, const void** h_where);
// device -> caller's host buffer (directly when it is page-locked, else cudaMemcpyAsync)
int small_d2h_to(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st);
bool device_can_read_host(const void* p);       // page-locked (or registered) host memory the device can address
int launch_sm_copy(void* dst, const void* src, size_t bytes, cudaStream_t st);

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
