// pcie_floor.cu -- the copy-only floor of the host-to-host batch path on 1..N GPUs of one box.
//
//   nvcc -O2 -o tools/build/pcie_floor tools/pcie_floor.cu
//   tools/build/pcie_floor [--max-gpus N] [--in-mb 838] [--out-mb 1061] [--chunk-mb 105] [--iters 4] [--quick]
//
// One process per GPU (forked before any CUDA call, like the ranks torchrun starts), started together
// through a barrier in shared memory.  Every test moves `in-mb` host->device and `out-mb` device->host
// per iteration and GPU, in `chunk-mb` pieces on two CUDA streams, exactly the bytes a C2 step of
// PipelinedCodec.round_trip moves (bench.py e2e: h2d_bytes_per_step / d2h_bytes_per_step), with no
// kernel at all.  The sweep crosses
//   GPU sets     {0} {0,1} {0,1,2,3} {0..7} plus pairs that probe shared uplinks ({0,2} {0,4} ...)
//   host memory  cudaMallocHost | cudaHostAlloc(portable) | THP-backed mmap + cudaHostRegister |
//                MAP_HUGETLB + cudaHostRegister (when the box has huge pages reserved)
//   direction    in only | out only | both at once
//   placement    unpinned | each rank's threads on a disjoint core set (first touch after pinning)
// and prints one JSON line per test: per-GPU and aggregate GB/s, and ms per "step".
// Development / measurement aid: nothing in the product imports it.
#include <cuda_runtime.h>
#include <errno.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>
#include <string>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "[rank %d] %s -> %s\n", g_rank, #x, cudaGetErrorString(e_));        \
            exit(3);                                                                            \
        }                                                                                       \
    } while (0)

static int g_rank = 0;

enum Alloc { A_MALLOCHOST = 0, A_PORTABLE = 1, A_THP = 2, A_HUGETLB = 3, A_COUNT = 4 };
static const char* alloc_name[A_COUNT] = {"cudaMallocHost", "cudaHostAlloc_portable", "thp_mmap_register", "hugetlb_register"};
enum Dir { D_IN = 1, D_OUT = 2, D_BOTH = 3 };

struct Test {
    unsigned gpu_mask;
    int alloc, dir, pin;      // pin: 0 none, 1 disjoint core sets in rank order, 2 reversed
    int chunk_mb;
};

struct Shared {
    pthread_barrier_t bar;
    double ms[16];            // per rank: wall ms of the last test
    int ok[16][A_COUNT];      // which allocations worked on each rank
};

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void pin_to(int first, int count) {
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c = first; c < first + count; ++c) CPU_SET(c, &set);
    sched_setaffinity(0, sizeof(set), &set);
}

static void unpin(int ncpu) {
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c = 0; c < ncpu; ++c) CPU_SET(c, &set);
    sched_setaffinity(0, sizeof(set), &set);
}

struct HostBuf {
    void* p = nullptr;        // usable pointer
    void* raw = nullptr;      // mmap base (registered kinds)
    size_t raw_len = 0;
    int kind = -1;
};

static void host_free(HostBuf& b) {
    if (!b.p) return;
    if (b.kind <= A_PORTABLE) cudaFreeHost(b.p);
    else {
        cudaHostUnregister(b.p);
        munmap(b.raw, b.raw_len);
    }
    b = HostBuf();
}

static HostBuf host_buffer(int kind, size_t bytes) {
    HostBuf b;
    b.kind = kind;
    const size_t two_mb = 2u << 20;
    const size_t rounded = (bytes + two_mb - 1) / two_mb * two_mb;
    switch (kind) {
        case A_MALLOCHOST:
            if (cudaMallocHost(&b.p, bytes) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return b; }
            break;
        case A_PORTABLE:
            if (cudaHostAlloc(&b.p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return b; }
            break;
        case A_THP: {
            b.raw_len = rounded + two_mb;
            b.raw = mmap(nullptr, b.raw_len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (b.raw == MAP_FAILED) return HostBuf();
            b.p = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(b.raw) + two_mb - 1) / two_mb * two_mb);
            madvise(b.p, rounded, MADV_HUGEPAGE);
            memset(b.p, 1, rounded);
            if (cudaHostRegister(b.p, rounded, cudaHostRegisterDefault) != cudaSuccess) {
                cudaGetLastError();
                munmap(b.raw, b.raw_len);
                return HostBuf();
            }
            break;
        }
        case A_HUGETLB: {
            b.raw_len = rounded;
            b.raw = mmap(nullptr, rounded, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
            if (b.raw == MAP_FAILED) return HostBuf();
            memset(b.raw, 1, rounded);
            if (cudaHostRegister(b.raw, rounded, cudaHostRegisterDefault) != cudaSuccess) {
                cudaGetLastError();
                munmap(b.raw, b.raw_len);
                return HostBuf();
            }
            b.p = b.raw;
            break;
        }
    }
    if (b.p && kind <= A_PORTABLE) memset(b.p, 1, bytes);
    return b;
}

int main(int argc, char** argv) {
    int max_gpus = 8, iters = 4, quick = 0, chunk_mb = 105;
    size_t in_mb = 838, out_mb = 1061;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--max-gpus") && i + 1 < argc) max_gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--in-mb") && i + 1 < argc) in_mb = atol(argv[++i]);
        else if (!strcmp(argv[i], "--out-mb") && i + 1 < argc) out_mb = atol(argv[++i]);
        else if (!strcmp(argv[i], "--chunk-mb") && i + 1 < argc) chunk_mb = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quick")) quick = 1;
    }
    // GPU count without creating a context in the parent (children must fork before CUDA starts)
    int n_gpus = 0;
    {
        FILE* f = popen("nvidia-smi -L 2>/dev/null | wc -l", "r");
        if (f) {
            if (fscanf(f, "%d", &n_gpus) != 1) n_gpus = 0;
            pclose(f);
        }
    }
    if (n_gpus <= 0) {
        fprintf(stderr, "no GPUs\n");
        return 2;
    }
    if (n_gpus > max_gpus) n_gpus = max_gpus;
    if (n_gpus > 16) n_gpus = 16;
    const int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN);

    // ---- the test list (identical in every process) ----
    std::vector<Test> tests;
    std::vector<unsigned> sets;
    for (int n = 1; n <= n_gpus; n *= 2) sets.push_back((1u << n) - 1u);
    if (n_gpus >= 4) { sets.push_back(0x5); sets.push_back(0x9); }            // {0,2} {0,3}
    if (n_gpus >= 8) { sets.push_back(0x11); sets.push_back(0x81); sets.push_back(0x55); sets.push_back(0xF0); }   // {0,4} {0,7} {0,2,4,6} {4..7}
    for (unsigned m : sets)                                                    // baseline memory, all directions
        for (int d : {D_BOTH, D_IN, D_OUT}) tests.push_back(Test{m, A_MALLOCHOST, d, 0, chunk_mb});
    const unsigned full = (1u << n_gpus) - 1u;
    std::vector<unsigned> big_sets;
    big_sets.push_back(full);
    if (n_gpus >= 4) big_sets.push_back((1u << (n_gpus / 2)) - 1u);
    if (!quick) {
        // (grouped by memory kind and placement: a rank keeps one pair of host buffers at a time)
        for (unsigned m : big_sets)
            for (int c : {16, 420}) tests.push_back(Test{m, A_MALLOCHOST, D_BOTH, 0, c});
        for (int a = 1; a < A_COUNT; ++a)
            for (unsigned m : big_sets) tests.push_back(Test{m, a, D_BOTH, 0, chunk_mb});
        for (int a : {A_MALLOCHOST, A_THP})
            for (int pin : {1, 2})
                for (unsigned m : big_sets) tests.push_back(Test{m, a, D_BOTH, pin, chunk_mb});
    }

    Shared* sh = static_cast<Shared*>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    memset(sh, 0, sizeof(*sh));
    pthread_barrierattr_t ba;
    pthread_barrierattr_init(&ba);
    pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &ba, n_gpus + 1);

    std::vector<pid_t> kids;
    for (int r = 0; r < n_gpus; ++r) {
        pid_t pid = fork();
        if (pid == 0) {
            g_rank = r;
            CK(cudaSetDevice(r));
            CK(cudaFree(0));
            const size_t in_b = in_mb << 20, out_b = out_mb << 20;
            void *d_in, *d_out;
            CK(cudaMalloc(&d_in, in_b));
            CK(cudaMalloc(&d_out, out_b));
            CK(cudaMemset(d_out, 7, out_b));
            cudaStream_t s_in, s_out;
            CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
            // one pair of host buffers at a time, first touched under the placement of the test that uses it
            HostBuf hb_in, hb_out;
            int cur_a = -1, cur_pin = -1;
            auto ensure = [&](int a, int pin) {
                if (a == cur_a && pin == cur_pin) return hb_in.p && hb_out.p;
                host_free(hb_in);
                host_free(hb_out);
                if (pin) {
                    const int per = ncpu / n_gpus > 0 ? ncpu / n_gpus : 1;
                    const int slot = pin == 1 ? r : n_gpus - 1 - r;
                    pin_to((slot * per) % ncpu, per);
                } else {
                    unpin(ncpu);
                }
                hb_in = host_buffer(a, in_b);
                hb_out = host_buffer(a, out_b);
                cur_a = a;
                cur_pin = pin;
                return hb_in.p && hb_out.p;
            };
            for (size_t t = 0; t < tests.size(); ++t) {
                const Test& T = tests[t];
                const bool mine = (T.gpu_mask >> r) & 1u;
                bool ok = true;
                if (mine) ok = ensure(T.alloc, T.pin);
                sh->ok[r][T.alloc] = ok ? 1 : 0;
                if (mine && ok) {
                    if (T.pin) {
                        const int per = ncpu / n_gpus > 0 ? ncpu / n_gpus : 1;
                        const int slot = T.pin == 1 ? r : n_gpus - 1 - r;
                        pin_to((slot * per) % ncpu, per);
                    } else {
                        unpin(ncpu);
                    }
                }
                const size_t chunk = (size_t)T.chunk_mb << 20;
                auto one_iter = [&]() {
                    if (T.dir & D_IN)
                        for (size_t o = 0; o < in_b; o += chunk)
                            CK(cudaMemcpyAsync((char*)d_in + o, (char*)hb_in.p + o, chunk < in_b - o ? chunk : in_b - o, cudaMemcpyHostToDevice, s_in));
                    if (T.dir & D_OUT)
                        for (size_t o = 0; o < out_b; o += chunk)
                            CK(cudaMemcpyAsync((char*)hb_out.p + o, (char*)d_out + o, chunk < out_b - o ? chunk : out_b - o, cudaMemcpyDeviceToHost, s_out));
                };
                if (mine && ok) {       // warm-up
                    one_iter();
                    CK(cudaStreamSynchronize(s_in));
                    CK(cudaStreamSynchronize(s_out));
                }
                pthread_barrier_wait(&sh->bar);
                const double t0 = now_ms();
                if (mine && ok) {
                    for (int it = 0; it < iters; ++it) one_iter();
                    CK(cudaStreamSynchronize(s_in));
                    CK(cudaStreamSynchronize(s_out));
                }
                sh->ms[r] = mine && ok ? (now_ms() - t0) / iters : 0.0;
                pthread_barrier_wait(&sh->bar);
            }
            _exit(0);
        }
        kids.push_back(pid);
    }
    printf("{\"info\": {\"gpus\": %d, \"cpus\": %d, \"in_mb\": %zu, \"out_mb\": %zu, \"iters\": %d}}\n", n_gpus, ncpu, in_mb, out_mb, iters);
    for (size_t t = 0; t < tests.size(); ++t) {
        const Test& T = tests[t];
        pthread_barrier_wait(&sh->bar);
        pthread_barrier_wait(&sh->bar);
        double worst = 0;
        int n = 0;
        bool ok = true;
        std::string per = "[";
        for (int r = 0; r < n_gpus; ++r)
            if ((T.gpu_mask >> r) & 1u) {
                ok = ok && sh->ok[r][T.alloc];
                worst = sh->ms[r] > worst ? sh->ms[r] : worst;
                char b[32];
                snprintf(b, sizeof(b), "%s%.1f", n ? ", " : "", sh->ms[r]);
                per += b;
                ++n;
            }
        per += "]";
        const double mb = ((T.dir & D_IN) ? in_mb : 0) + ((T.dir & D_OUT) ? out_mb : 0);
        printf("{\"gpus\": \"0x%x\", \"n\": %d, \"alloc\": \"%s\", \"dir\": \"%s\", \"pin\": %d, \"chunk_mb\": %d, \"ok\": %s, "
               "\"ms_per_step\": %.2f, \"per_rank_ms\": %s, \"aggregate_gbs\": %.1f, \"per_gpu_gbs\": %.1f}\n",
               T.gpu_mask, n, alloc_name[T.alloc], T.dir == D_BOTH ? "both" : (T.dir == D_IN ? "in" : "out"), T.pin, T.chunk_mb,
               ok ? "true" : "false", worst, per.c_str(), ok && worst > 0 ? n * mb * 1.048576 / worst : 0.0,
               ok && worst > 0 ? mb * 1.048576 / worst : 0.0);
        fflush(stdout);
    }
    for (pid_t k : kids) {
        int st;
        waitpid(k, &st, 0);
    }
    return 0;
}
