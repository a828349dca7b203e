// hic_wavelet.cu -- wavelet ("HIC") mode transform stage: K9 (colour + 3-level db1 + sub-band
// quantisation + threshold + whole-sub-band zigzag, fused), K10 (its inverse) and layout converters.
//
// Reference: compression.wavelet_compression / wavelet_decompression (compression.py:59-100),
// quantization.subband_quantize / subband_invert_quantize (quantization.py:60-77), transform.threshold
// (transform.py:227-239), codec.wavelet_encode's zigzag of every whole sub-band (codec.py:123-126) and
// codec.wavelet_decode_pull_subbands (codec.py:166-179).  The transform itself is PyWavelets'
// wavedec2 / waverec2 with "db1" in float64; its arithmetic (operation order of the down/up-sampling
// convolutions, symmetric extension) is restated from the published algorithm in
// oracle/pywt_standin.py -- PARITY UNPINNED against the real package (not installable here).
//
// Data layout.  Output of K9 / input of K10 is the "flat stream" the entropy stage consumes
// (hic_layout_flat): per image and channel (lum, cr, cb) the ten sub-bands
// [cA3, cH3, cV3, cD3, cH2, cV2, cD2, cH1, cV1, cD1], each zigzagged as a whole, concatenated, int16,
// padded to a multiple of 64 elements.  3 samples per pixel: 3 B/pixel in, 6 B/pixel out.
//
// One 8x8 pixel tile maps to exactly one cA3 sample, so a thread owns a tile and runs all three
// levels in registers (float64, non-contracted, PyWavelets' operation order).  A CTA owns 128x128
// pixels; its coefficients are staged in shared memory and written along the anti-diagonals of each
// sub-band region, which are contiguous runs of the zigzag order.
#include "hic_core.cuh"
#include "hic_runtime.cuh"
#include "hic_wavelet_common.cuh"

namespace hic {
namespace wv {

constexpr int REGION = 128;                  // pixels per CTA side
constexpr int TPS = REGION / 8;              // tiles per side = 16
constexpr int THREADS = TPS * TPS;           // 256
constexpr int RGB_PITCH = REGION * 3;        // bytes per staged row
constexpr double C = 0x1.6a09e667f3bcdp-1;   // PyWavelets' db1 coefficient 7.071067811865475244e-01 as a double

struct Geom {
    hic_wavelet_geometry w;
    int64_t chan_elems;          // 64 * blocks per channel
};

// shared-memory staging of one channel's coefficients for a 128x128 region: level-1 bands (3 x 64x64),
// level-2 (3 x 32x32), level-3 and cA (4 x 16x16)
__host__ __device__ constexpr int stage_off(int band) {
    return band >= 7 ? (band - 7) * 4096 : (band >= 4 ? 12288 + (band - 4) * 1024 : 15360 + band * 256);
}
__host__ __device__ constexpr int band_level(int band) { return band >= 7 ? 1 : (band >= 4 ? 2 : 3); }

__device__ __forceinline__ void haar_pair(double even, double odd, double& a, double& d) {
    const double ce = dmul(C, even), co = dmul(C, odd);
    a = dadd(co, ce);              // (c * x[2k+1]) + (c * x[2k])
    d = dadd(-co, ce);             // (-c * x[2k+1]) + (c * x[2k])
}

// quantise one coefficient: true division, np.round (half to even), |v| < 5 -> 0 (quantization.py:60-69,
// transform.py:227-239 at the default settings)
__device__ __forceinline__ int16_t quant(double v, double div) {
    const int q = __double2int_rn(div == 1.0 ? v : ddiv(v, div));
    return (int16_t)((q < 5 && q > -5) ? 0 : q);
}

// One level on an N x N register tile (N = 8, 4, 2).  in: N x N (row major, pitch N); `ry`, `rx`: valid
// rows / columns of `in` (symmetric extension: a missing odd partner repeats the even sample).
// aa: (N/2) x (N/2) approximations; the three detail bands go to the staging area through `emit`.
template <int N, typename Emit>
__device__ __forceinline__ void level(const double* in, int ry, int rx, double* aa, Emit emit) {
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        if (2 * k >= ry) break;
        const bool has_odd_row = 2 * k + 1 < ry;
        double a0[N], d0[N];
#pragma unroll
        for (int x = 0; x < N; ++x) {
            const double e = in[(2 * k) * N + x], o = has_odd_row ? in[(2 * k + 1) * N + x] : e;
            haar_pair(e, o, a0[x], d0[x]);          // axis 0 first (pywt.dwtn)
        }
#pragma unroll
        for (int j = 0; j < N / 2; ++j) {
            if (2 * j >= rx) break;
            const bool has_odd_col = 2 * j + 1 < rx;
            double v_aa, v_ad, v_da, v_dd;
            haar_pair(a0[2 * j], has_odd_col ? a0[2 * j + 1] : a0[2 * j], v_aa, v_ad);
            haar_pair(d0[2 * j], has_odd_col ? d0[2 * j + 1] : d0[2 * j], v_da, v_dd);
            aa[k * (N / 2) + j] = v_aa;
            emit(k, j, v_da, v_ad, v_dd);           // cH = 'da', cV = 'ad', cD = 'dd'
        }
    }
}

__global__ void __launch_bounds__(THREADS)
forward_kernel(const uint8_t* __restrict__ rgb, Geom g, int16_t* __restrict__ flat) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* s_rgb = smem_raw;                                        // REGION rows x RGB_PITCH bytes
    int16_t* s_out = reinterpret_cast<int16_t*>(smem_raw + REGION * RGB_PITCH);
    const int h = g.w.h, w = g.w.w;
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * REGION, y0 = blockIdx.y * REGION;
    const int tid = threadIdx.x;
    const uint8_t* src = rgb + (size_t)img * h * w * 3;

    // ---- stage the RGB region (coalesced 16-byte loads when rows are 16-byte aligned) ----
    const int rows = min(REGION, h - y0), cols = min(REGION, w - x0);
    const int row_bytes = cols * 3;
    const bool vec = ((3 * w) % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (row_bytes % 16 == 0);
    if (vec) {
        const int per_row = row_bytes / 16;
        for (int i = tid; i < rows * per_row; i += THREADS) {
            const int r = i / per_row, c = i - r * per_row;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)(y0 + r) * w + x0) * 3) + c);
            *reinterpret_cast<uint4*>(s_rgb + r * RGB_PITCH + 16 * c) = v;
        }
    } else {
        for (int i = tid; i < rows * row_bytes; i += THREADS) {
            const int r = i / row_bytes, c = i - r * row_bytes;
            s_rgb[r * RGB_PITCH + c] = __ldg(src + ((size_t)(y0 + r) * w + x0) * 3 + c);
        }
    }
    __syncthreads();

    // ---- colour conversion of this thread's 8x8 tile, all three channels packed as bytes ----
    const int ty = tid / TPS, tx = tid % TPS;
    const int ry0 = max(0, min(8, rows - 8 * ty)), rx0 = max(0, min(8, cols - 8 * tx));
    const bool live = ry0 > 0 && rx0 > 0;
    uint32_t ycc[3][16];                   // [channel][row * 2 + half]: four samples per word
    if (live) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(s_rgb + (8 * ty + r) * RGB_PITCH + 24 * tx);
            uint32_t wv[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) wv[j] = p[j];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const uint32_t w0 = wv[3 * hf], w1 = wv[3 * hf + 1], w2 = wv[3 * hf + 2];
                const uint32_t px[4] = {w0, __funnelshift_r(w0, w1, 24), __funnelshift_r(w1, w2, 16), w2 >> 8};
                uint32_t py = 0, pcr = 0, pcb = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int yy, cr, cb;
                    rgb_to_ycrcb((int)(px[k] & 0xFF), (int)((px[k] >> 8) & 0xFF), (int)((px[k] >> 16) & 0xFF), yy, cr, cb);
                    py |= (uint32_t)yy << (8 * k);
                    pcr |= (uint32_t)cr << (8 * k);
                    pcb |= (uint32_t)cb << (8 * k);
                }
                ycc[0][2 * r + hf] = py;
                ycc[1][2 * r + hf] = pcr;
                ycc[2][2 * r + hf] = pcb;
            }
        }
    }
    const int ry1 = (ry0 + 1) >> 1, rx1 = (rx0 + 1) >> 1, ry2 = (ry1 + 1) >> 1, rx2 = (rx1 + 1) >> 1;

#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        if (live) {
            double a1[16], a2[4], a3[1];
            {
                double px[64];
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    px[i] = (double)((int)((ycc[ch][i >> 2] >> (8 * (i & 3))) & 0xFF) - 256);      // compression.py:70
                level<8>(px, ry0, rx0, a1, [&](int k, int j, double da, double ad, double dd) {
                    const int o = (4 * ty + k) * 64 + 4 * tx + j;
                    s_out[stage_off(7) + o] = quant(da, 5.0);
                    s_out[stage_off(8) + o] = quant(ad, 5.0);
                    s_out[stage_off(9) + o] = quant(dd, 5.0);
                });
            }
            level<4>(a1, ry1, rx1, a2, [&](int k, int j, double da, double ad, double dd) {
                const int o = (2 * ty + k) * 32 + 2 * tx + j;
                s_out[stage_off(4) + o] = quant(da, 2.0);
                s_out[stage_off(5) + o] = quant(ad, 2.0);
                s_out[stage_off(6) + o] = quant(dd, 2.0);
            });
            level<2>(a2, ry2, rx2, a3, [&](int, int, double da, double ad, double dd) {
                const int o = ty * 16 + tx;
                s_out[stage_off(1) + o] = quant(da, 1.0);
                s_out[stage_off(2) + o] = quant(ad, 1.0);
                s_out[stage_off(3) + o] = quant(dd, 1.0);
            });
            s_out[stage_off(0) + ty * 16 + tx] = quant(a3[0], 1.0);
        }
        __syncthreads();
        // ---- write-out along anti-diagonals: a diagonal of a region is a contiguous zigzag run ----
        int16_t* dst = flat + ((size_t)img * 3 + ch) * g.chan_elems;
        const int warp = tid >> 5, lane = tid & 31;
        for (int band = 0; band < 10; ++band) {
            const int lvl = band_level(band);
            const int R = REGION >> lvl;                       // region side in this band
            const int hb = g.w.lh[lvl], wb = g.w.lw[lvl];
            const int Y0 = y0 >> lvl, X0 = x0 >> lvl;
            const int vr = min(R, hb - Y0), vc = min(R, wb - X0);      // valid part of the region
            if (vr <= 0 || vc <= 0) continue;
            const int16_t* st = s_out + stage_off(band);
            int16_t* out = dst + g.w.band_off[band];
            for (int dl = warp; dl < vr + vc - 1; dl += THREADS / 32) {
                const int d = Y0 + X0 + dl;
                const int64_t base = diag_start(d, hb, wb);
                const int y_lo = max(0, d - (wb - 1)), y_hi = min(d, hb - 1);
                const int i_lo = max(0, dl - (vc - 1)), i_hi = min(dl, vr - 1);
                for (int i = i_lo + lane; i <= i_hi; i += 32) {
                    const int Y = Y0 + i;
                    const int64_t pos = base + ((d & 1) ? (y_hi - Y) : (Y - y_lo));
                    out[pos] = st[i * R + (dl - i)];
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K10: inverse.  Reference wavelet_decompression (compression.py:88-100): detail bands times
// (i*i + 1), pywt.waverec2 (idwtn undoes the LAST axis first; x[2k] = (c a[k]) + (c d[k]),
// x[2k+1] = (c a[k]) + (-c d[k])), + 256, astype(uint8) (truncate toward zero, wrap), cvtColor.
// Only shapes that are multiples of 8 (the reference's own decoder assumes exact doubling,
// codec.py:182-189).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ihaar_pair(double a, double d, double& even, double& odd) {
    const double ca = dmul(C, a), cd = dmul(C, d);
    even = dadd(ca, cd);
    odd = dadd(ca, -cd);
}

// (N x N) approximations + three (N x N) detail bands -> (2N x 2N)
template <int N>
__device__ __forceinline__ void ilevel(const double* aa, const double* da, const double* ad, const double* dd, double* out) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double a0[2 * N], d0[2 * N];         // axis 1 first: rows of the axis-0 approximation / detail
#pragma unroll
        for (int j = 0; j < N; ++j) {
            ihaar_pair(aa[k * N + j], ad[k * N + j], a0[2 * j], a0[2 * j + 1]);
            ihaar_pair(da[k * N + j], dd[k * N + j], d0[2 * j], d0[2 * j + 1]);
        }
#pragma unroll
        for (int x = 0; x < 2 * N; ++x) ihaar_pair(a0[x], d0[x], out[(2 * k) * 2 * N + x], out[(2 * k + 1) * 2 * N + x]);
    }
}

__global__ void __launch_bounds__(THREADS)
inverse_kernel(const int16_t* __restrict__ flat, Geom g, uint8_t* __restrict__ rgb) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* s_rgb = smem_raw;
    int16_t* s_in = reinterpret_cast<int16_t*>(smem_raw + REGION * RGB_PITCH);
    const int h = g.w.h, w = g.w.w;
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * REGION, y0 = blockIdx.y * REGION;
    const int tid = threadIdx.x;
    const int rows = min(REGION, h - y0), cols = min(REGION, w - x0);
    const int ty = tid / TPS, tx = tid % TPS;
    const bool live = 8 * ty < rows && 8 * tx < cols;
    uint32_t ycc[3][16];

#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        // ---- gather this region's coefficients along the anti-diagonals ----
        const int16_t* srcp = flat + ((size_t)img * 3 + ch) * g.chan_elems;
        const int warp = tid >> 5, lane = tid & 31;
        for (int band = 0; band < 10; ++band) {
            const int lvl = band_level(band);
            const int R = REGION >> lvl;
            const int hb = g.w.lh[lvl], wb = g.w.lw[lvl];
            const int Y0 = y0 >> lvl, X0 = x0 >> lvl;
            const int vr = min(R, hb - Y0), vc = min(R, wb - X0);
            if (vr <= 0 || vc <= 0) continue;
            int16_t* st = s_in + stage_off(band);
            const int16_t* in = srcp + g.w.band_off[band];
            for (int dl = warp; dl < vr + vc - 1; dl += THREADS / 32) {
                const int d = Y0 + X0 + dl;
                const int64_t base = diag_start(d, hb, wb);
                const int y_lo = max(0, d - (wb - 1)), y_hi = min(d, hb - 1);
                const int i_lo = max(0, dl - (vc - 1)), i_hi = min(dl, vr - 1);
                for (int i = i_lo + lane; i <= i_hi; i += 32) {
                    const int Y = Y0 + i;
                    const int64_t pos = base + ((d & 1) ? (y_hi - Y) : (Y - y_lo));
                    st[i * R + (dl - i)] = __ldg(in + pos);
                }
            }
        }
        __syncthreads();
        if (live) {
            // quantization.subband_invert_quantize: detail level i (0 = coarsest) times i*i + 1
            double a3[1], h3[1], v3[1], d3[1], a2[4];
            const int o3 = ty * 16 + tx;
            a3[0] = (double)s_in[stage_off(0) + o3];
            h3[0] = (double)s_in[stage_off(1) + o3];
            v3[0] = (double)s_in[stage_off(2) + o3];
            d3[0] = (double)s_in[stage_off(3) + o3];
            ilevel<1>(a3, h3, v3, d3, a2);
            double hh[16], vv[16], dd[16], a1[16];
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int o = (2 * ty + k) * 32 + 2 * tx + j;
                    hh[k * 2 + j] = (double)(2 * (int)s_in[stage_off(4) + o]);
                    vv[k * 2 + j] = (double)(2 * (int)s_in[stage_off(5) + o]);
                    dd[k * 2 + j] = (double)(2 * (int)s_in[stage_off(6) + o]);
                }
            ilevel<2>(a2, hh, vv, dd, a1);
            double px[64];
            {
                double h1[16], v1[16], d1[16];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int o = (4 * ty + k) * 64 + 4 * tx + j;
                        h1[k * 4 + j] = (double)(5 * (int)s_in[stage_off(7) + o]);
                        v1[k * 4 + j] = (double)(5 * (int)s_in[stage_off(8) + o]);
                        d1[k * 4 + j] = (double)(5 * (int)s_in[stage_off(9) + o]);
                    }
                ilevel<4>(a1, h1, v1, d1, px);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t wd = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) wd |= (uint32_t)wrap_u8(dadd(px[4 * i + k], 256.0)) << (8 * k);
                ycc[ch][i] = wd;
            }
        }
        __syncthreads();
    }
    // ---- YCrCb -> RGB into the staging area, then coalesced row writes ----
    if (live) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint8_t* o = s_rgb + (8 * ty + r) * RGB_PITCH + 24 * tx;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int sh = 8 * (c & 3), wi = 2 * r + (c >> 2);
                int rr, gg, bb;
                ycrcb_to_rgb((int)((ycc[0][wi] >> sh) & 0xFF), (int)((ycc[1][wi] >> sh) & 0xFF), (int)((ycc[2][wi] >> sh) & 0xFF), rr, gg, bb);
                o[3 * c] = (uint8_t)rr;
                o[3 * c + 1] = (uint8_t)gg;
                o[3 * c + 2] = (uint8_t)bb;
            }
        }
    }
    __syncthreads();
    uint8_t* dstp = rgb + (size_t)img * h * w * 3;
    const int row_bytes = cols * 3;
    const bool vec = ((3 * w) % 16 == 0) && ((reinterpret_cast<uintptr_t>(dstp) & 15) == 0) && (row_bytes % 16 == 0);
    if (vec) {
        const int per_row = row_bytes / 16;
        for (int i = tid; i < rows * per_row; i += THREADS) {
            const int r = i / per_row, c = i - r * per_row;
            *(reinterpret_cast<uint4*>(dstp + ((size_t)(y0 + r) * w + x0) * 3) + c) =
                *reinterpret_cast<const uint4*>(s_rgb + r * RGB_PITCH + 16 * c);
        }
    } else {
        for (int i = tid; i < rows * row_bytes; i += THREADS) {
            const int r = i / row_bytes, c = i - r * row_bytes;
            dstp[((size_t)(y0 + r) * w + x0) * 3 + c] = s_rgb[r * RGB_PITCH + c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// layout converters at the CompressedImage edge: flat zigzag stream <-> ten raster int32 sub-bands
// (concatenated per channel in band order).  One thread per coefficient.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
flat_to_bands_kernel(const int16_t* __restrict__ flat, Geom g, int n, int32_t* __restrict__ bands) {
    const int64_t per = g.w.len;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per * 3 * n) return;
    const int64_t cs = gid / per;              // image * 3 + channel
    const int64_t e = gid - cs * per;          // raster position inside the channel's concatenated bands
    int band = 9;
    while (e < g.w.band_off[band]) --band;
    const int lvl = band_level(band);
    const int hb = g.w.lh[lvl], wb = g.w.lw[lvl];
    const int64_t r = e - g.w.band_off[band];
    const int y = (int)(r / wb), x = (int)(r - (int64_t)y * wb);
    bands[gid] = (int32_t)flat[cs * g.chan_elems + g.w.band_off[band] + zigzag_pos(y, x, hb, wb)];
}

__global__ void __launch_bounds__(256)
bands_to_flat_kernel(const int32_t* __restrict__ bands, Geom g, int n, int16_t* __restrict__ flat) {
    const int64_t per = g.w.len;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per * 3 * n) return;
    const int64_t cs = gid / per;
    const int64_t e = gid - cs * per;
    int band = 9;
    while (e < g.w.band_off[band]) --band;
    const int lvl = band_level(band);
    const int hb = g.w.lh[lvl], wb = g.w.lw[lvl];
    const int64_t r = e - g.w.band_off[band];
    const int y = (int)(r / wb), x = (int)(r - (int64_t)y * wb);
    flat[cs * g.chan_elems + g.w.band_off[band] + zigzag_pos(y, x, hb, wb)] = (int16_t)bands[gid];
}

static int geometry_of(int h, int w, hic_wavelet_geometry* g) {
    HIC_REQUIRE(g != nullptr, "geometry output is NULL");
    HIC_REQUIRE(h >= 1 && w >= 1 && h <= 65536 && w <= 65536, "image must be 1..65536 on a side (got %dx%d)", h, w);
    g->h = h;
    g->w = w;
    g->lh[0] = h;
    g->lw[0] = w;
    for (int l = 1; l <= 3; ++l) {           // pywt.dwt_coeff_len for db1, symmetric mode: ceil(n / 2)
        g->lh[l] = (g->lh[l - 1] + 1) / 2;
        g->lw[l] = (g->lw[l - 1] + 1) / 2;
    }
    int64_t off = 0;
    for (int b = 0; b < 10; ++b) {
        const int lvl = band_level(b);
        g->band_off[b] = off;
        off += (int64_t)g->lh[lvl] * g->lw[lvl];
    }
    g->len = off;
    return HIC_OK;
}

static inline unsigned ceil_div_u(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }
constexpr size_t SMEM_BYTES = (size_t)REGION * RGB_PITCH + 16384 * sizeof(int16_t);

static int ensure_attrs() {
    static bool done[64] = {false};
    int dev = 0;
    HIC_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return HIC_OK;
    HIC_CUDA(cudaFuncSetAttribute(forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    HIC_CUDA(cudaFuncSetAttribute(inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    if (dev < 64) done[dev] = true;
    return HIC_OK;
}

}  // namespace wv
}  // namespace hic

extern "C" {

int hic_wavelet_geometry_of(int32_t h, int32_t w, hic_wavelet_geometry* out) { return hic::wv::geometry_of(h, w, out); }

int hic_wavelet_forward(const uint8_t* d_rgb, int32_t n, int32_t h, int32_t w, int16_t* d_flat, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_rgb && d_flat, "NULL device pointer");
    HIC_REQUIRE(n >= 1 && n <= 65535, "batch size must be in 1..65535 (got %d)", n);
    wv::Geom g;
    int rc = wv::geometry_of(h, w, &g.w);
    if (rc) return rc;
    g.chan_elems = 64 * ((g.w.len + 63) / 64);
    rc = wv::ensure_attrs();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    // the tail of every channel's last 64-element block is padding the entropy stage never reads past len
    dim3 grid(wv::ceil_div_u(w, wv::REGION), wv::ceil_div_u(h, wv::REGION), n);
    HIC_LAUNCH("wavelet_forward_kernel", st, wv::forward_kernel<<<grid, wv::THREADS, wv::SMEM_BYTES, st>>>(d_rgb, g, d_flat));
    return HIC_OK;
}

int hic_wavelet_inverse(const int16_t* d_flat, int32_t n, int32_t h, int32_t w, uint8_t* d_rgb_out, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_rgb_out && d_flat, "NULL device pointer");
    HIC_REQUIRE(n >= 1 && n <= 65535, "batch size must be in 1..65535 (got %d)", n);
    HIC_REQUIRE(h % 8 == 0 && w % 8 == 0, "wavelet decode needs multiples of 8 (got %dx%d), as the reference's decoder does", h, w);
    wv::Geom g;
    int rc = wv::geometry_of(h, w, &g.w);
    if (rc) return rc;
    g.chan_elems = 64 * ((g.w.len + 63) / 64);
    rc = wv::ensure_attrs();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    dim3 grid(wv::ceil_div_u(w, wv::REGION), wv::ceil_div_u(h, wv::REGION), n);
    HIC_LAUNCH("wavelet_inverse_kernel", st, wv::inverse_kernel<<<grid, wv::THREADS, wv::SMEM_BYTES, st>>>(d_flat, g, d_rgb_out));
    return HIC_OK;
}

int hic_wavelet_flat_to_bands(const int16_t* d_flat, int32_t n, int32_t h, int32_t w, int32_t* d_bands, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_flat && d_bands, "NULL device pointer");
    HIC_REQUIRE(n >= 1, "batch size must be positive");
    wv::Geom g;
    int rc = wv::geometry_of(h, w, &g.w);
    if (rc) return rc;
    g.chan_elems = 64 * ((g.w.len + 63) / 64);
    cudaStream_t st = as_stream(stream);
    const int64_t items = g.w.len * 3 * n;
    HIC_LAUNCH("wavelet_flat_to_bands_kernel", st, wv::flat_to_bands_kernel<<<wv::ceil_div_u(items, 256), 256, 0, st>>>(d_flat, g, n, d_bands));
    return HIC_OK;
}

int hic_wavelet_bands_to_flat(const int32_t* d_bands, int32_t n, int32_t h, int32_t w, int16_t* d_flat, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_flat && d_bands, "NULL device pointer");
    HIC_REQUIRE(n >= 1, "batch size must be positive");
    wv::Geom g;
    int rc = wv::geometry_of(h, w, &g.w);
    if (rc) return rc;
    g.chan_elems = 64 * ((g.w.len + 63) / 64);
    cudaStream_t st = as_stream(stream);
    const int64_t items = g.w.len * 3 * n;
    HIC_CUDA(cudaMemsetAsync(d_flat, 0, (size_t)g.chan_elems * 3 * n * sizeof(int16_t), st));
    HIC_LAUNCH("wavelet_bands_to_flat_kernel", st, wv::bands_to_flat_kernel<<<wv::ceil_div_u(items, 256), 256, 0, st>>>(d_bands, g, n, d_flat));
    return HIC_OK;
}

}  // extern "C"
