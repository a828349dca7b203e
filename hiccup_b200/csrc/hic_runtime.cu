// hic_runtime.cu -- device/stream/memory plumbing of the C ABI (include/hiccup_b200.h).
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>
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

namespace {
struct ProfSpan {
    const char* name;
    cudaEvent_t a, b;
    cudaStream_t stream;
};
bool g_prof_on = false;
std::vector<ProfSpan> g_spans;
std::vector<cudaEvent_t> g_free_events;
std::mutex g_prof_mu;
cudaEvent_t take_event() {
    if (!g_free_events.empty()) {
        cudaEvent_t e = g_free_events.back();
        g_free_events.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

namespace {
thread_local cudaEvent_t t_open_end = nullptr;      // end event of the span this host thread has open
}

void prof_begin(const char* name, cudaStream_t st) {
    t_open_end = nullptr;
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    ProfSpan sp{name, take_event(), take_event(), st};
    cudaEventRecord(sp.a, st);
    g_spans.push_back(sp);
    t_open_end = sp.b;
}

void prof_end(cudaStream_t st) {
    if (!t_open_end) return;
    cudaEventRecord(t_open_end, st);
    t_open_end = nullptr;
}

// ---- small transfers through the SMs ----------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
sm_copy_kernel(char* __restrict__ dst, const char* __restrict__ src, size_t bytes, int vec16) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (vec16) {
        const size_t n16 = bytes / 16;
        uint4* d = reinterpret_cast<uint4*>(dst);
        const uint4* s = reinterpret_cast<const uint4*>(src);
        for (size_t i = tid; i < n16; i += step) d[i] = s[i];
        done = n16 * 16;
    }
    for (size_t i = done + tid; i < bytes; i += step) dst[i] = src[i];
}

}  // namespace

bool device_can_read_host(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr;
}

int launch_sm_copy(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return HIC_OK;
    const int vec16 = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
    const size_t items = vec16 ? bytes / 16 + 16 : bytes;
    const unsigned grid = (unsigned)std::min<size_t>(64, (items + 255) / 256);
    sm_copy_kernel<<<grid, 256, 0, st>>>(static_cast<char*>(dst), static_cast<const char*>(src), bytes, vec16);
    HIC_CHECK_LAUNCH("sm_copy_kernel");
    return HIC_OK;
}

namespace {
int stage_reserve(SmallXfer& x, size_t bytes, char** out) {
    const size_t need = (x.used + 15) / 16 * 16 + bytes;
    if (need > x.cap) {
        if (x.used != 0) return hic::fail(HIC_ERR_CAPACITY, "small-transfer staging of %zu bytes is in use; %zu more needed", x.cap, bytes);
        if (x.h_stage) cudaFreeHost(x.h_stage);
        x.h_stage = nullptr;
        x.cap = std::max<size_t>(need + need / 2, 1u << 20);
        HIC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&x.h_stage), x.cap));
    }
    x.used = (x.used + 15) / 16 * 16;
    *out = x.h_stage + x.used;
    x.used += bytes;
    return HIC_OK;
}
}  // namespace

void SmallXfer::destroy() {
    if (h_stage) cudaFreeHost(h_stage);
    h_stage = nullptr;
    cap = used = 0;
}

int small_h2d(SmallXfer& x, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return HIC_OK;
    if (bytes > SMALL_XFER_MAX) {
        HIC_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return HIC_OK;
    }
    if (device_can_read_host(h_src)) return launch_sm_copy(d_dst, h_src, bytes, st);
    char* stage = nullptr;
    if (x.cap - std::min(x.cap, (x.used + 15) / 16 * 16) < bytes && x.used != 0) {      // no room left this call: the plain way
        HIC_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return HIC_OK;
    }
    int rc = stage_reserve(x, bytes, &stage);
    if (rc) return rc;
    memcpy(stage, h_src, bytes);
    return launch_sm_copy(d_dst, stage, bytes, st);
}

int small_d2h(SmallXfer& x, const void* d_src, size_t bytes, cudaStream_t st, const void** h_where) {
    char* stage = nullptr;
    int rc = stage_reserve(x, bytes, &stage);
    if (rc) return rc;
    *h_where = stage;
    return launch_sm_copy(stage, d_src, bytes, st);
}

int small_d2h_to(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return HIC_OK;
    if (bytes <= SMALL_XFER_MAX && device_can_read_host(h_dst)) return launch_sm_copy(h_dst, d_src, bytes, st);
    HIC_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    return HIC_OK;
}

}  // namespace hic

extern "C" {

int hic_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(hic::g_prof_mu);
    hic::g_prof_on = on != 0;
    return HIC_OK;
}

int hic_profile_report(char* buf, size_t buflen) {
    HIC_REQUIRE(buf != nullptr && buflen > 2, "buf is NULL");
    std::lock_guard<std::mutex> lock(hic::g_prof_mu);
    std::map<std::string, std::pair<double, int>> agg;
    for (auto& sp : hic::g_spans) {
        cudaEventSynchronize(sp.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            auto& e = agg[sp.name];
            e.first += ms;
            e.second += 1;
        }
        hic::g_free_events.push_back(sp.a);
        hic::g_free_events.push_back(sp.b);
    }
    hic::g_spans.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char item[256];
        snprintf(item, sizeof(item), "%s\"%s\": [%.6f, %d]", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
        out += item;
        first = false;
    }
    out += "}";
    if (out.size() + 1 > buflen) return hic::fail(HIC_ERR_CAPACITY, "profile report needs %zu bytes", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return HIC_OK;
}

int hic_profile_timeline(char* buf, size_t buflen) {
    HIC_REQUIRE(buf != nullptr && buflen > 2, "buf is NULL");
    std::lock_guard<std::mutex> lock(hic::g_prof_mu);
    std::string out = "[";
    std::map<cudaStream_t, int> ids;
    bool first = true;
    for (auto& sp : hic::g_spans) {
        cudaEventSynchronize(sp.b);
        float t0 = 0.f, t1 = 0.f;
        if (cudaEventElapsedTime(&t0, hic::g_spans.front().a, sp.a) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&t1, hic::g_spans.front().a, sp.b) != cudaSuccess) continue;
        const int id = ids.emplace(sp.stream, (int)ids.size()).first->second;
        char item[256];
        snprintf(item, sizeof(item), "%s[\"%s\", %d, %.4f, %.4f]", first ? "" : ", ", sp.name, id, t0, t1);
        out += item;
        first = false;
    }
    out += "]";
    if (out.size() + 1 > buflen) return hic::fail(HIC_ERR_CAPACITY, "profile timeline needs %zu bytes", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return HIC_OK;
}

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

int hic_get_device(int* device) {
    HIC_REQUIRE(device != nullptr, "device is NULL");
    HIC_CUDA(cudaGetDevice(device));
    return HIC_OK;
}

int hic_set_blocking_sync(int on) {
    // how host threads wait for the device: yielding the core (blocking) instead of spinning (opt-in: the
    // pipelined batch path keeps one host thread per slot waiting most of the time)
    HIC_CUDA(cudaSetDeviceFlags(on ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto));
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

int hic_host_register(void* h_ptr, size_t bytes) {
    HIC_REQUIRE(h_ptr != nullptr && bytes > 0, "nothing to register");
    HIC_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterPortable));
    return HIC_OK;
}

int hic_host_unregister(void* h_ptr) {
    HIC_CUDA(cudaHostUnregister(h_ptr));
    return HIC_OK;
}

int hic_ticket_take(void* h_counter, int64_t count, int64_t* first) {
    HIC_REQUIRE(h_counter != nullptr && first != nullptr, "NULL argument");
    HIC_REQUIRE((reinterpret_cast<uintptr_t>(h_counter) & 7) == 0, "the counter must be 8-byte aligned");
    *first = __atomic_fetch_add(static_cast<int64_t*>(h_counter), count, __ATOMIC_ACQ_REL);
    return HIC_OK;
}

// Bulk copies (image batches, compressed payloads) are single asynchronous copy-engine transfers.  A copy
// engine serves its queue strictly in order (tools/ce_fifo.py: a 64 KB copy on another stream waits the
// full ~2 ms behind a queued 105 MB copy, also when that copy is enqueued as 4 MB pieces).  What helped the
// pipelined batch path was taking the SMALL transfers off the engines (SmallXfer above: 27.9 -> 25.3 ms
// per 1024-image batch, tools/pipe_one.py).  Two other remedies were measured and dropped: pacing the
// bulk pieces from the host so that other copies can slip in between (30.7 ms), and sending the 20 MB
// compressed payloads of a chunk through the SMs as well (28.4 ms).
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
