// hic_core.cuh -- arithmetic shared by the sm_100a kernels and the host-side check harness.
//
// Everything here is `__host__ __device__` so that the exact same arithmetic the kernels run can be
// exercised on the CPU (tests/cpu_harness) before any GPU time is spent.  Reference citations are
// into /root/reference (nhomble/hiccup).
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define HIC_HD __host__ __device__ __forceinline__
#else
#define HIC_HD inline
#endif

namespace hic {

// ------------------------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------------------------
// Annex-K tables, natural (row-major) order -- reference hiccup/quantization.py:14-37.
#define HIC_LUM_TABLE { \
    16, 11, 10, 16, 24, 40, 51, 61,   12, 12, 14, 19, 26, 58, 60, 55, \
    14, 13, 16, 24, 40, 57, 69, 56,   14, 17, 22, 29, 51, 87, 80, 62, \
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, \
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 }
#define HIC_CHROMA_TABLE { \
    17, 18, 24, 47, 99, 99, 99, 99,   18, 21, 26, 66, 99, 99, 99, 99, \
    24, 26, 56, 99, 99, 99, 99, 99,   47, 66, 99, 99, 99, 99, 99, 99, \
    99, 99, 99, 99, 99, 99, 99, 99,   99, 99, 99, 99, 99, 99, 99, 99, \
    99, 99, 99, 99, 99, 99, 99, 99,   99, 99, 99, 99, 99, 99, 99, 99 }

// Scan order of reference transform._zigzag_indices (transform.py:106-124): anti-diagonals d = x+y,
// even d with y ascending, odd d with y descending (the transpose of the standard JPEG scan).
// Entry k is the natural index 8*y + x read at scan position k.
#define HIC_ZIGZAG8 { \
    0, 8, 1, 2, 9, 16, 24, 17, 10, 3, 4, 11, 18, 25, 32, 40, 33, 26, 19, 12, 5, 6, 13, 20, 27, 34, 41, 48, \
    56, 49, 42, 35, 28, 21, 14, 7, 15, 22, 29, 36, 43, 50, 57, 58, 51, 44, 37, 30, 23, 31, 38, 45, 52, \
    59, 60, 53, 46, 39, 47, 54, 61, 62, 55, 63 }

// ------------------------------------------------------------------------------------------------
// float64 helpers that are never contracted into FMAs (the reference's scipy/numpy builds are
// plain x86-64: separate multiply and add, each rounded).  Host builds use -ffp-contract=off.
// ------------------------------------------------------------------------------------------------
HIC_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
HIC_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
HIC_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
HIC_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

// ------------------------------------------------------------------------------------------------
// scipy.fftpack.dct / idct (ducc0) for N = 8, operation for operation in float64.
//
// ducc0 T_dcst23<double>::exec: type 2 = pre-butterflies, backward real FFT of length 8 run as a
// radix-2 pass (l1=1, ido=4) then a radix-4 pass (l1=2, ido=1), post-twiddle; type 3 = the mirror.
// Twiddles are ducc0's UnityRoots<double> values (two small tables multiplied in double), which
// differ from correctly rounded cosines by an ulp; they are embedded as hex literals and checked
// against the oracle's run-time derivation in tests/test_oracle_dct.py.
// Reference call sites: hiccup/transform.py:80,82 (dct2) and :99,101 (idct2).
// ------------------------------------------------------------------------------------------------
#define HIC_TW0 0x1.f6297cff75cb0p-1   /* cos(1*pi/16) */
#define HIC_TW1 0x1.d906bcf328d46p-1   /* cos(2*pi/16) */
#define HIC_TW2 0x1.a9b66290ea1a3p-1   /* cos(3*pi/16) */
#define HIC_TW3 0x1.6a09e667f3bccp-1   /* cos(4*pi/16), one ulp below the rounded value */
#define HIC_TW4 0x1.1c73b39ae68c8p-1   /* cos(5*pi/16) */
#define HIC_TW5 0x1.87de2a6aea963p-2   /* cos(6*pi/16) */
#define HIC_TW6 0x1.8f8b83c69a60ap-3   /* cos(7*pi/16) */
#define HIC_WR  0x1.6a09e667f3bccp-1   /* Re exp(2 pi i / 8) */
#define HIC_WI  0x1.6a09e667f3bcdp-1   /* Im exp(2 pi i / 8) */

HIC_HD void ducc_dct2_8(double* c) {
    const double c0 = dmul(c[0], 2.0), c7 = dmul(c[7], 2.0);
    const double c2 = dsub(c[2], c[1]), c1 = dadd(c[2], c[1]);
    const double c4 = dsub(c[4], c[3]), c3 = dadd(c[4], c[3]);
    const double c6 = dsub(c[6], c[5]), c5 = dadd(c[6], c[5]);
    // radix-2 backward pass
    double a[8];
    a[0] = dadd(c0, c7);
    a[4] = dsub(c0, c7);
    a[3] = dmul(2.0, c3);
    a[7] = dmul(-2.0, c4);
    a[1] = dadd(c1, c5);
    const double tr2 = dsub(c1, c5);
    const double ti2 = dadd(c2, c6);
    a[2] = dsub(c2, c6);
    a[6] = dadd(dmul(HIC_WR, ti2), dmul(HIC_WI, tr2));
    a[5] = dsub(dmul(HIC_WR, tr2), dmul(HIC_WI, ti2));
    // radix-4 backward pass
    double r[8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double t2 = dadd(a[4 * k], a[4 * k + 3]);
        const double t1 = dsub(a[4 * k], a[4 * k + 3]);
        const double t3 = dmul(2.0, a[4 * k + 1]);
        const double t4 = dmul(2.0, a[4 * k + 2]);
        r[k] = dadd(t2, t3);
        r[k + 4] = dsub(t2, t3);
        r[k + 6] = dadd(t1, t4);
        r[k + 2] = dsub(t1, t4);
    }
    c[0] = r[0];
    {
        const double t1 = dadd(dmul(HIC_TW0, r[7]), dmul(HIC_TW6, r[1]));
        const double t2 = dsub(dmul(HIC_TW0, r[1]), dmul(HIC_TW6, r[7]));
        c[1] = dmul(0.5, dadd(t1, t2));
        c[7] = dmul(0.5, dsub(t1, t2));
    }
    {
        const double t1 = dadd(dmul(HIC_TW1, r[6]), dmul(HIC_TW5, r[2]));
        const double t2 = dsub(dmul(HIC_TW1, r[2]), dmul(HIC_TW5, r[6]));
        c[2] = dmul(0.5, dadd(t1, t2));
        c[6] = dmul(0.5, dsub(t1, t2));
    }
    {
        const double t1 = dadd(dmul(HIC_TW2, r[5]), dmul(HIC_TW4, r[3]));
        const double t2 = dsub(dmul(HIC_TW2, r[3]), dmul(HIC_TW4, r[5]));
        c[3] = dmul(0.5, dadd(t1, t2));
        c[5] = dmul(0.5, dsub(t1, t2));
    }
    c[4] = dmul(r[4], HIC_TW3);
}

HIC_HD void ducc_dct3_8(double* c) {
    {
        const double t1 = dadd(c[1], c[7]), t2 = dsub(c[1], c[7]);
        c[1] = dadd(dmul(HIC_TW0, t2), dmul(HIC_TW6, t1));
        c[7] = dsub(dmul(HIC_TW0, t1), dmul(HIC_TW6, t2));
    }
    {
        const double t1 = dadd(c[2], c[6]), t2 = dsub(c[2], c[6]);
        c[2] = dadd(dmul(HIC_TW1, t2), dmul(HIC_TW5, t1));
        c[6] = dsub(dmul(HIC_TW1, t1), dmul(HIC_TW5, t2));
    }
    {
        const double t1 = dadd(c[3], c[5]), t2 = dsub(c[3], c[5]);
        c[3] = dadd(dmul(HIC_TW2, t2), dmul(HIC_TW4, t1));
        c[5] = dsub(dmul(HIC_TW2, t1), dmul(HIC_TW4, t2));
    }
    c[4] = dmul(c[4], dmul(2.0, HIC_TW3));
    // radix-4 forward pass (l1=2, ido=1)
    double h[8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double p0 = c[k], p1 = c[k + 2], p2 = c[k + 4], p3 = c[k + 6];
        const double tr1 = dadd(p3, p1);
        h[2 + 4 * k] = dsub(p3, p1);
        const double tr2 = dadd(p0, p2);
        h[1 + 4 * k] = dsub(p0, p2);
        h[0 + 4 * k] = dadd(tr2, tr1);
        h[3 + 4 * k] = dsub(tr2, tr1);
    }
    // radix-2 forward pass (l1=1, ido=4)
    double r[8];
    r[0] = dadd(h[0], h[4]);
    r[7] = dsub(h[0], h[4]);
    r[4] = -h[7];
    r[3] = h[3];
    const double tr2 = dadd(dmul(HIC_WR, h[5]), dmul(HIC_WI, h[6]));
    const double ti2 = dsub(dmul(HIC_WR, h[6]), dmul(HIC_WI, h[5]));
    r[1] = dadd(h[1], tr2);
    r[5] = dsub(h[1], tr2);
    r[2] = dadd(ti2, h[2]);
    r[6] = dsub(ti2, h[2]);
    c[0] = r[0];
    c[7] = r[7];
    c[1] = dsub(r[1], r[2]);
    c[2] = dadd(r[1], r[2]);
    c[3] = dsub(r[3], r[4]);
    c[4] = dadd(r[3], r[4]);
    c[5] = dsub(r[5], r[6]);
    c[6] = dadd(r[5], r[6]);
}

// round half to even, as np.round (reference quantization.py:80-81)
HIC_HD int32_t round_half_even(double v) {
#ifdef __CUDA_ARCH__
    return __double2int_rn(v);
#else
    return (int32_t)nearbyint(v);
#endif
}

// One quantised coefficient exactly as the reference computes it: rows then columns through
// ducc0's DCT (transform.py:67-84), true division by the table entry, np.round, int32.
// px: the 8x8 block after the -128 shift and zero padding, row major.  (u, v) = (row, col).
HIC_HD int32_t exact_quantised_coef(const int16_t* px, int u, int v, int q) {
    double col[8];
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        double row[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = (double)px[8 * i + j];
        ducc_dct2_8(row);
        col[i] = row[v];
    }
    ducc_dct2_8(col);
    return round_half_even(ddiv(col[u], (double)q));
}

// The whole quantised block exactly as the reference computes it (16 eight-point transforms).
// px: x - 128, zero padded, row major; q: the table, row major; out: natural order.
HIC_HD void exact_quantised_block(const int16_t* px, const int* q, int32_t* out) {
    double a[64];
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        double row[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = (double)px[8 * i + j];
        ducc_dct2_8(row);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[8 * i + j] = row[j];
    }
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        double col[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) col[i] = a[8 * i + j];
        ducc_dct2_8(col);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[8 * i + j] = round_half_even(ddiv(col[i], (double)q[8 * i + j]));
    }
}

// The whole decoded block exactly as the reference computes it.  cq: coef * table, natural order;
// out: the float64 samples before the uint8 cast, row major.
HIC_HD void exact_decoded_block(const int32_t* cq, double* out) {
    double a[64];
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        double row[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = (double)cq[8 * i + j];
        ducc_dct3_8(row);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[8 * i + j] = row[j];
    }
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        double col[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) col[i] = a[8 * i + j];
        ducc_dct3_8(col);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[8 * i + j] = dadd(ddiv(col[i], 256.0), 128.0);
    }
}

// One decoded sample exactly as the reference computes it (transform.py:169-179, 87-103):
// coef*table, idct rows then columns, /256, +128, astype(uint8) = truncate toward zero, wrap.
// cq: the dequantised 8x8 block (coef * table), row major.  Returns the float64 value before the
// uint8 cast so callers can count near-integer cases.
HIC_HD double exact_decoded_sample(const int32_t* cq, int y, int x) {
    double col[8];
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        double row[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) row[j] = (double)cq[8 * i + j];
        ducc_dct3_8(row);
        col[i] = row[x];
    }
    ducc_dct3_8(col);
    return dadd(ddiv(col[y], 256.0), 128.0);
}

HIC_HD uint8_t wrap_u8(double v) {            // numpy float64 -> uint8 cast on x86-64
    return (uint8_t)((long long)v & 0xFF);    // (long long) truncates toward zero
}

// ------------------------------------------------------------------------------------------------
// colour conversion: OpenCV's 14-bit fixed point (reference compression.py:21,56)
// ------------------------------------------------------------------------------------------------
HIC_HD void rgb_to_ycrcb(int r, int g, int b, int& y, int& cr, int& cb) {
    y = (4899 * r + 9617 * g + 1868 * b + 8192) >> 14;
    cr = ((r - y) * 11682 + (128 << 14) + 8192) >> 14;
    cb = ((b - y) * 9241 + (128 << 14) + 8192) >> 14;
    cr = cr < 0 ? 0 : (cr > 255 ? 255 : cr);
    cb = cb < 0 ? 0 : (cb > 255 ? 255 : cb);
}

HIC_HD int clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

HIC_HD void ycrcb_to_rgb(int y, int cr, int cb, int& r, int& g, int& b) {
    cr -= 128;
    cb -= 128;
    r = clamp_u8(y + ((cr * 22987 + 8192) >> 14));
    g = clamp_u8(y + ((cb * -5636 + cr * -11698 + 8192) >> 14));
    b = clamp_u8(y + ((cb * 29049 + 8192) >> 14));
}

// BORDER_REFLECT_101 index, then clamped (tiles may overhang far past a small image; those
// positions are masked out downstream, the clamp only keeps the address legal)
HIC_HD int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// ------------------------------------------------------------------------------------------------
// fast float32 8-point transforms (Arai-Agui-Nakajima factorisation: 5 multiplies, 29 adds).
// Forward: out[k] = S_k / g_k with S_k = sum_n x_n cos((2n+1) k pi / 16), g_0 = 1,
// g_k = 1 / (2 cos(k pi / 16)).  Inverse: with in[k] = X_k * h_k, h_0 = 1, h_k = 2 cos(k pi / 16),
// out[n] = X_0 + 2 sum_k X_k cos(pi k (2n+1) / 16)  (scipy's unnormalised DCT-III).
// The scale factors are folded into the quantisation tables by the callers.
// ------------------------------------------------------------------------------------------------
template <typename T>
HIC_HD void aan_forward8(T& d0, T& d1, T& d2, T& d3, T& d4, T& d5, T& d6, T& d7) {
    const T t0 = d0 + d7, t7 = d0 - d7, t1 = d1 + d6, t6 = d1 - d6;
    const T t2 = d2 + d5, t5 = d2 - d5, t3 = d3 + d4, t4 = d3 - d4;
    const T e10 = t0 + t3, e13 = t0 - t3, e11 = t1 + t2, e12 = t1 - t2;
    d0 = e10 + e11;
    d4 = e10 - e11;
    const T z1 = (e12 + e13) * T(0.70710678118654752440);
    d2 = e13 + z1;
    d6 = e13 - z1;
    const T o10 = t4 + t5, o11 = t5 + t6, o12 = t6 + t7;
    const T z5 = (o10 - o12) * T(0.38268343236508977173);
    const T z2 = T(0.54119610014619698440) * o10 + z5;
    const T z4 = T(1.30656296487637652786) * o12 + z5;
    const T z3 = o11 * T(0.70710678118654752440);
    const T z11 = t7 + z3, z13 = t7 - z3;
    d5 = z13 + z2;
    d3 = z13 - z2;
    d1 = z11 + z4;
    d7 = z11 - z4;
}

template <typename T>
HIC_HD void aan_inverse8(T& d0, T& d1, T& d2, T& d3, T& d4, T& d5, T& d6, T& d7) {
    const T e10 = d0 + d4, e11 = d0 - d4, e13 = d2 + d6;
    const T e12 = (d2 - d6) * T(1.41421356237309504880) - e13;
    const T t0 = e10 + e13, t3 = e10 - e13, t1 = e11 + e12, t2 = e11 - e12;
    const T z13 = d5 + d3, z10 = d5 - d3, z11 = d1 + d7, z12 = d1 - d7;
    const T t7 = z11 + z13;
    const T o11 = (z11 - z13) * T(1.41421356237309504880);
    const T z5 = (z10 + z12) * T(1.84775906502257351225);
    const T o10 = z5 - z12 * T(1.08239220029239396880);
    const T o12 = z5 - z10 * T(2.61312592975275305571);
    const T t6 = o12 - t7, t5 = o11 - t6, t4 = o10 - t5;
    d0 = t0 + t7;
    d7 = t0 - t7;
    d1 = t1 + t6;
    d6 = t1 - t6;
    d2 = t2 + t5;
    d5 = t2 - t5;
    d3 = t3 + t4;
    d4 = t3 - t4;
}

// forward scale g_k and inverse prescale h_k (see above)
HIC_HD double aan_g(int k) { return k == 0 ? 1.0 : 1.0 / (2.0 * cos(k * 3.14159265358979323846 / 16.0)); }
HIC_HD double aan_h(int k) { return k == 0 ? 1.0 : 2.0 * cos(k * 3.14159265358979323846 / 16.0); }

// Safety factor of the near-tie band: a float32 quantised value v = C/q is trusted when its
// distance to the nearest half-integer exceeds HIC_TIE_KAPPA * 2^-24 * 4 * E / q, E = sum |x| over
// the block.  (A first-order bound for the two AAN passes is ~14; tests/cpu_harness measures the
// observed maximum, see DESIGN.md.)
#define HIC_TIE_KAPPA 16.0
// Same idea for the decode transform: a float32 sample p is trusted when its distance to the nearest
// integer (where the reference's uint8 truncation steps) exceeds HIC_INV_KAPPA * 2^-24 * 4 * S / 256,
// S = sum |coef * q| over the block.
#define HIC_INV_KAPPA 16.0

}  // namespace hic
