// hic_f32x2.cuh -- packed two-lane float32 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2, PTX
// add/mul/fma.rn.f32x2, sm_100+).  One instruction does two independent float32 operations on a
// 64-bit register pair; the DCT butterflies and the quantiser run two rows (or two columns) at once.
#pragma once
#include <stdint.h>

namespace hic {

struct f2 {
    unsigned long long v;
    __device__ __forceinline__ f2() {}
    __device__ __forceinline__ f2(float s) { asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(s)); }
    __device__ __forceinline__ f2(double s) : f2((float)s) {}
    __device__ __forceinline__ f2(float a, float b) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b)); }
    __device__ __forceinline__ float lo() const {
        float a, b;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
        return a;
    }
    __device__ __forceinline__ float hi() const {
        float a, b;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
        return b;
    }
};
__device__ __forceinline__ f2 operator+(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 operator-(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 operator*(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}

// the generic name the transform templates of hic_core.cuh use
__device__ __forceinline__ f2 eo_fma(f2 a, f2 b, f2 c) { return fma2(a, b, c); }

}  // namespace hic
