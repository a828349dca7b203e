// CPU check harness: runs the kernels' shared __host__ __device__ arithmetic (hic_core.cuh) on the
// host so it can be compared with the oracle without a GPU.  TEST INFRASTRUCTURE ONLY -- it is not
// linked into libhiccup_b200.so and nothing under hiccup_b200/ loads it.
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC harness.cpp -o libhic_cpu_harness.so
#include <string.h>
#include <stdlib.h>
#include "../../hiccup_b200/csrc/hic_core.cuh"
#include "../../hiccup_b200/csrc/hic_replay.cuh"
#include "../../hiccup_b200/csrc/hic_dct_bound.h"

using namespace hic;

static const int LUM[64] = HIC_LUM_TABLE;
static const int CHROMA[64] = HIC_CHROMA_TABLE;
static const uint8_t ZZ[64] = HIC_ZIGZAG8;

extern "C" {

void hx_ducc_dct2(double* x, int n) { for (int i = 0; i < n; ++i) ducc_dct2_8(x + 8 * i); }
void hx_ducc_dct3(double* x, int n) { for (int i = 0; i < n; ++i) ducc_dct3_8(x + 8 * i); }
int hx_exact_coef(const int16_t* px, int u, int v, int q) { return exact_quantised_coef(px, u, v, q); }
double hx_exact_sample(const int32_t* cq, int y, int x) { return exact_decoded_sample(cq, y, x); }

// The rigorous error-bound tables the library builds at initialisation (csrc/hic_dct_bound.h): natural order.
void hx_dct_bounds(double* kappa_fwd, double* w_inv, double* fwd_scale, double* inv_prescale) {
    const DctBounds b = dct_bounds();
    for (int i = 0; i < 64; ++i) {
        kappa_fwd[i] = b.kappa_fwd[i];
        w_inv[i] = b.w_inv[i];
    }
    for (int k = 0; k < 8; ++k) {
        fwd_scale[k] = eo_forward_scale(k);
        inv_prescale[k] = eo_inverse_prescale(k);
    }
}

// float32 forward path of K1's transform_pair, restated on the host with the same primitives.
// px: [nb][64] int16 (x-128, row major); out: [nb][64] int16 in scan order (float32 result, before
// fix-up); mask: [nb] near-tie masks; ratio: [nb] max over AC coefficients of
// |v32 - v64| * q / (kappa(u, v) * 4 * 2^-24 * E): the observed float32 error as a fraction of the rigorous
// bound of that coefficient (must stay below 1; the kernel's band is the bound times HIC_BAND_MARGIN).
void hx_forward_blocks(const int16_t* px, int nb, int kind, int16_t* out, uint64_t* mask, double* ratio) {
    const int* q = kind == 0 ? LUM : CHROMA;
    const DctBounds bounds = dct_bounds();
    float rq[64], kq[64];
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            rq[8 * u + v] = (float)(4.0 / (eo_forward_scale(u) * eo_forward_scale(v) * q[8 * u + v]));
            kq[8 * u + v] = (float)(bounds.kappa_fwd[8 * u + v] * HIC_BAND_MARGIN * 4.0 / 16777216.0 / q[8 * u + v]);
        }
    const float MAGIC = 12582912.0f;
    for (int b = 0; b < nb; ++b) {
        float v[64];
        float abs_sum = 0.f;
        for (int i = 0; i < 64; ++i) { v[i] = (float)px[64 * b + i]; abs_sum += fabsf(v[i]); }
        for (int r = 0; r < 8; ++r)
            eo_forward8(v[8*r], v[8*r+1], v[8*r+2], v[8*r+3], v[8*r+4], v[8*r+5], v[8*r+6], v[8*r+7]);
        for (int c = 0; c < 8; ++c)
            eo_forward8(v[c], v[8+c], v[16+c], v[24+c], v[32+c], v[40+c], v[48+c], v[56+c]);
        // float64 truth for the ratio
        double truth[64];
        {
            double a[64];
            for (int i = 0; i < 8; ++i) {
                double row[8];
                for (int j = 0; j < 8; ++j) row[j] = px[64 * b + 8 * i + j];
                ducc_dct2_8(row);
                for (int j = 0; j < 8; ++j) a[8 * i + j] = row[j];
            }
            for (int j = 0; j < 8; ++j) {
                double col[8];
                for (int i = 0; i < 8; ++i) col[i] = a[8 * i + j];
                ducc_dct2_8(col);
                for (int i = 0; i < 8; ++i) truth[8 * i + j] = col[i];
            }
        }
        uint64_t m = 0;
        double worst = 0.0;
        for (int k = 0; k < 64; ++k) {
            const int nat = ZZ[k];
            const float t = fmaf(v[nat], rq[nat], MAGIC);
            int32_t bits;
            memcpy(&bits, &t, 4);
            out[64 * b + k] = (int16_t)(bits & 0xFFFF);
            if (k != 0) {
                const float d = fmaf(v[nat], rq[nat], MAGIC - t);
                if (fmaf(abs_sum, kq[nat], fabsf(d)) >= 0.499999f) m |= (1ull << k);
                if (abs_sum > 0.f) {
                    const double v32 = (double)v[nat] * (double)rq[nat];
                    const double v64 = truth[nat] / q[nat];
                    const double r = fabs(v32 - v64) * q[nat] / (bounds.kappa_fwd[nat] * 4.0 / 16777216.0 * abs_sum);
                    if (r > worst) worst = r;
                }
            }
        }
        mask[b] = m;
        ratio[b] = worst;
    }
}

// float32 inverse path of K7's inverse_block.  coef: [nb][64] int16 scan order; out [nb][64] uint8
// (float32 result before fix-up); mask: samples within the near-integer band; ratio: max over
// samples of |p32 - p64| / (2^-24 * 4 / 256 * sum_k w_k |coef_k q_k|): the observed error as a fraction of the
// rigorous bound (must stay below 1); exact: [nb][64] uint8 from the float64 emulation.
void hx_inverse_blocks(const int16_t* coef, int nb, int kind, uint8_t* out, uint64_t* mask, double* ratio, uint8_t* exact) {
    const int* q = kind == 0 ? LUM : CHROMA;
    const DctBounds bounds = dct_bounds();
    float dq[64], qw[64];
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            dq[8 * u + v] = (float)(q[8 * u + v] * eo_inverse_prescale(u) * eo_inverse_prescale(v) / 256.0);
            qw[8 * u + v] = (float)(q[8 * u + v] * bounds.w_inv[8 * u + v] * HIC_BAND_MARGIN);
        }
    for (int b = 0; b < nb; ++b) {
        float v[64];
        int32_t cq[64];
        float S = 0.f;
        double S_exact = 0.0;
        bool ac = false;
        for (int k = 0; k < 64; ++k) {
            const int c = coef[64 * b + k];
            v[ZZ[k]] = (float)c * dq[ZZ[k]];
            cq[ZZ[k]] = c * q[ZZ[k]];
            S = fmaf(fabsf((float)c), qw[ZZ[k]], S);
            S_exact += fabs((double)c) * q[ZZ[k]] * bounds.w_inv[ZZ[k]];
            if (k && c) ac = true;
        }
        for (int r = 0; r < 8; ++r)
            eo_inverse8(v[8*r], v[8*r+1], v[8*r+2], v[8*r+3], v[8*r+4], v[8*r+5], v[8*r+6], v[8*r+7]);
        for (int c = 0; c < 8; ++c)
            eo_inverse8(v[c], v[8+c], v[16+c], v[24+c], v[32+c], v[40+c], v[48+c], v[56+c]);
        const float band = (float)(4.0 / 16777216.0 / 256.0) * S + 3.0517578125e-5f;
        uint64_t m = 0;
        double worst = 0.0;
        for (int i = 0; i < 64; ++i) {
            const float p = v[i] + 128.f;
            out[64 * b + i] = (uint8_t)((int)p & 0xFF);
            const double p64 = exact_decoded_sample(cq, i >> 3, i & 7);
            exact[64 * b + i] = wrap_u8(p64);
            if (ac && fabsf(p - rintf(p)) <= band) m |= (1ull << i);
            if (S_exact > 0.0) {
                // (the final + 128 rounds to 2^-15 at most: taken off before the comparison, as the band adds it back)
                const double err = fabs((double)p - p64) - 1.52587890625e-5;
                const double r = err / (4.0 / 16777216.0 / 256.0 * S_exact);
                if (r > worst) worst = r;
            }
        }
        mask[b] = m;
        ratio[b] = worst;
    }
}

void hx_colour(const uint8_t* rgb, int n, uint8_t* ycrcb) {
    for (int i = 0; i < n; ++i) {
        int y, cr, cb;
        rgb_to_ycrcb(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], y, cr, cb);
        ycrcb[3 * i] = (uint8_t)y; ycrcb[3 * i + 1] = (uint8_t)cr; ycrcb[3 * i + 2] = (uint8_t)cb;
    }
}

void hx_colour_inv(const uint8_t* ycrcb, int n, uint8_t* rgb) {
    for (int i = 0; i < n; ++i) {
        int r, g, b;
        ycrcb_to_rgb(ycrcb[3 * i], ycrcb[3 * i + 1], ycrcb[3 * i + 2], r, g, b);
        rgb[3 * i] = (uint8_t)r; rgb[3 * i + 1] = (uint8_t)g; rgb[3 * i + 2] = (uint8_t)b;
    }
}

// The device Huffman builder's packed heapq replay (csrc/hic_replay.cuh) on the host.  freqs: leaf
// frequencies in first-occurrence order (their sum below 2^18, 2 <= n <= 8192); par: 2 n parent links out
// (bit 15 = left child).  mode = 1 | 2: one or two heap levels per step (both are compiled into the library).
int hx_replay_narrow(const uint32_t* freqs, int n, int mode, uint16_t* par) {
    if (n < 2 || n > REPLAY_NARROW_MAX_LEAVES) return 1;
    unsigned long long total = 0;
    for (int i = 0; i < n; ++i) total += freqs[i];
    if (total >= REPLAY_NARROW_TOTAL) return 2;
    const int slots = (n + 4 + 15) / 16 * 16;
    uint32_t* slot = static_cast<uint32_t*>(aligned_alloc(64, sizeof(uint32_t) * slots));
    for (int i = 0; i < slots; ++i) slot[i] = 0xDEADBEEFu;        // stale words the prefetches may read
    for (int i = 0; i < n; ++i) slot[i + 1] = (freqs[i] << REPLAY_ID_BITS) | (uint32_t)i;
    for (int i = 0; i < 2 * n; ++i) par[i] = 0;
    if (mode == 2) replay_narrow<2>(slot, n, slots, par);
    else replay_narrow<1>(slot, n, slots, par);
    free(slot);
    return 0;
}

}  // extern "C"
