// pipe_rates.cu -- issue rates of the integer / packed instructions K1 is built from, on this GPU
// (development aid: which of dp4a / dp2a / vsadu4 / prmt / imad / i2f run at full rate?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/pipe_rates tools/scratch/pipe_rates.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(uint32_t* out, uint32_t seed) {
    uint32_t v[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
        v[c] = seed * (threadIdx.x + 1) + c * 0x9E3779B9u;
        f[c] = (float)(threadIdx.x + c);
    }
    const uint32_t k = seed | 0x01010101u;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) v[c] = __dp4a(v[c], k, v[c]);
            if (OP == 1) v[c] = __dp2a_lo(k, v[c], v[c]);
            if (OP == 2) v[c] = v[c] * k + seed;                                  // IMAD
            if (OP == 3) v[c] = __byte_perm(v[c], k, 0x0462);                     // PRMT
            if (OP == 4) v[c] = __vsadu4(v[c], k) + v[c];                         // VABSDIFF4 (+ accumulate)
            if (OP == 5) v[c] = __funnelshift_r(v[c], k, 8);                      // SHF
            if (OP == 6) v[c] = (v[c] & k) ^ seed;                                // LOP3
            if (OP == 7) f[c] = fmaf(f[c], 1.0001f, 0.5f);                        // FFMA
            if (OP == 8) f[c] = (float)(int)(v[c] = v[c] + 1u) + f[c];            // I2F (+ IADD, FADD)
            if (OP == 9) v[c] = min(v[c] + k, 0x00FFFFFFu);                       // IADD + VIMNMX
            if (OP == 10) v[c] = v[c] + k;                                        // IADD3
            if (OP == 11) f[c] = fmaxf(fabsf(f[c]) - 0.25f, fmaxf(f[c], 0.125f)); // FADD + FMNMX (x2)
            if (OP == 12) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(*reinterpret_cast<unsigned long long*>(&v[c & ~1])) : "l"(0x3f8000003f800000ull));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc += v[c] + __float_as_uint(f[c]);
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

template <int OP>
static void run(const char* name, double ops_per_iter_chain) {
    uint32_t* d;
    cudaMalloc(&d, 4096);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int grid = sms * 8;
    rate_kernel<OP><<<grid, 256>>>(d, 3);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    rate_kernel<OP><<<grid, 256>>>(d, 5);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double lane_ops = (double)grid * 256 * ITERS * CHAINS * ops_per_iter_chain;
    const double per_clk_sm = lane_ops / (ms * 1e-3) / (clock_khz * 1e3) / sms;
    printf("%-34s %8.3f ms  %7.1f lane-ops/clk/SM  (%.2f warp-instr/clk/SMSP)\n", name, ms, per_clk_sm, per_clk_sm / 32 / 4);
    cudaFree(d);
}

int main() {
    run<0>("IDP.4A (dp4a)", 1);
    run<1>("IDP.2A (dp2a)", 1);
    run<2>("IMAD", 1);
    run<3>("PRMT", 1);
    run<4>("VABSDIFF4.ACC (vsadu4 + add)", 1);
    run<5>("SHF (funnelshift)", 1);
    run<6>("LOP3", 1);
    run<7>("FFMA", 1);
    run<8>("I2F + IADD + FADD", 1);
    run<9>("IADD + VIMNMX", 1);
    run<10>("IADD3", 1);
    run<11>("FADD + 2 FMNMX", 1);
    run<12>("FADD2 (f32x2; 2 flops)", 1);
    return 0;
}
