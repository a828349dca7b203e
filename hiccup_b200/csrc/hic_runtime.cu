// hic_runtime.cu -- device/stream/memory plumbing of the C ABI (include/hiccup_b200.h).
#include <string.h>
#include "hic_runtime.cuh"

namespace hic {

char* last_error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace hic

extern "C" {

int hic_version(void) { return 100; }

const char* hic_last_error(void) { return hic::last_error_buffer(); }

int hic_device_count(int* count) {
    HIC_REQUIRE(count != nullptr, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return hic::fail(HIC_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return HIC_OK;
}

int hic_set_device(int device) {
    HIC_CUDA(cudaSetDevice(device));
    return HIC_OK;
}

int hic_device_name(char* buf, size_t buflen) {
    HIC_REQUIRE(buf != nullptr && buflen > 0, "buf is NULL");
    int dev = 0;
    HIC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    HIC_CUDA(cudaGetDeviceProperties(&prop, dev));
    snprintf(buf, buflen, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    return HIC_OK;
}

int hic_malloc(void** d_ptr, size_t bytes) {
    HIC_REQUIRE(d_ptr != nullptr, "d_ptr is NULL");
    HIC_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return HIC_OK;
}

int hic_free(void* d_ptr) {
    HIC_CUDA(cudaFree(d_ptr));
    return HIC_OK;
}

int hic_host_alloc(void** h_ptr, size_t bytes) {
    HIC_REQUIRE(h_ptr != nullptr, "h_ptr is NULL");
    HIC_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 1));
    return HIC_OK;
}

int hic_host_free(void* h_ptr) {
    HIC_CUDA(cudaFreeHost(h_ptr));
    return HIC_OK;
}

int hic_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
    HIC_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, hic::as_stream(stream)));
    return HIC_OK;
}

int hic_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
    HIC_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, hic::as_stream(stream)));
    return HIC_OK;
}

int hic_memset(void* d_ptr, int value, size_t bytes, void* stream) {
    HIC_CUDA(cudaMemsetAsync(d_ptr, value, bytes, hic::as_stream(stream)));
    return HIC_OK;
}

int hic_stream_create(void** stream) {
    HIC_REQUIRE(stream != nullptr, "stream is NULL");
    cudaStream_t s;
    HIC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return HIC_OK;
}

int hic_stream_destroy(void* stream) {
    HIC_CUDA(cudaStreamDestroy(hic::as_stream(stream)));
    return HIC_OK;
}

int hic_stream_sync(void* stream) {
    HIC_CUDA(cudaStreamSynchronize(hic::as_stream(stream)));
    return HIC_OK;
}

}  // extern "C"
