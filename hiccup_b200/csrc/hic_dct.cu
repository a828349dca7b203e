// hic_dct.cu -- DCT-mode transform stage: K1 (fused colour + pyrDown + DCT + quantise + zigzag),
// the float64 tie fix-up, K7 (dequantise + IDCT) and K8 (pyrUp + colour), plus layout converters.
//
// Data layout in HBM (see include/hiccup_b200.h): RGB u8 interleaved in; coefficients as int16
// "zigzag blocks" (one 128-byte line per 8x8 block) out.  HBM-bound by design: 3 B/pixel read,
// 3 B/pixel written (1.5 samples/pixel x 2 B) = 6 algorithmic bytes per pixel for K1.
#include <mutex>
#include "hic_core.cuh"
#include "hic_runtime.cuh"

namespace hic {

// ------------------------------------------------------------------------------------------------
// constant tables (per device)
// ------------------------------------------------------------------------------------------------
struct DctTables {
    float rq[2][64];     // [0 = luminance, 1 = chroma][natural index]: 4 g_u g_v / q  (forward)
    float qf[2][64];     // q as float
    float dq[2][64];     // q h_u h_v / 256 (inverse)
    int qi[2][64];       // q
};
__constant__ DctTables c_tab;
__constant__ uint8_t c_zigzag[64] = HIC_ZIGZAG8;

static const int h_lum[64] = HIC_LUM_TABLE;
static const int h_chroma[64] = HIC_CHROMA_TABLE;
static const uint8_t h_zigzag[64] = HIC_ZIGZAG8;

static int ensure_tables() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    HIC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && done[dev]) return HIC_OK;
    DctTables t;
    for (int kind = 0; kind < 2; ++kind)
        for (int u = 0; u < 8; ++u)
            for (int v = 0; v < 8; ++v) {
                const int q = (kind == 0 ? h_lum : h_chroma)[8 * u + v];
                t.rq[kind][8 * u + v] = (float)(4.0 * aan_g(u) * aan_g(v) / q);
                t.qf[kind][8 * u + v] = (float)q;
                t.dq[kind][8 * u + v] = (float)(q * aan_h(u) * aan_h(v) / 256.0);
                t.qi[kind][8 * u + v] = q;
            }
    HIC_CUDA(cudaMemcpyToSymbol(c_tab, &t, sizeof(t)));
    if (dev < 64) done[dev] = true;
    return HIC_OK;
}

// ------------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------------
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

static int geometry_of(int h, int w, hic_dct_geometry* g) {
    HIC_REQUIRE(g != nullptr, "geometry output is NULL");
    HIC_REQUIRE(h >= 2 && w >= 2, "image must be at least 2x2 (got %dx%d)", h, w);
    HIC_REQUIRE(h <= 65536 && w <= 65536, "image larger than 65536 on a side (got %dx%d)", h, w);
    g->h = h;
    g->w = w;
    g->hc = h / 2;
    g->wc = w / 2;
    g->nby_l = ceil_div(h, 8);
    g->nbx_l = ceil_div(w, 8);
    g->nby_c = ceil_div(g->hc, 8);
    g->nbx_c = ceil_div(g->wc, 8);
    g->nb_l = (int64_t)g->nby_l * g->nbx_l;
    g->nb_c = (int64_t)g->nby_c * g->nbx_c;
    g->blocks_per_image = g->nb_l + 2 * g->nb_c;
    g->out_h = 2 * g->hc;
    g->out_w = 2 * g->wc;
    return HIC_OK;
}

// ------------------------------------------------------------------------------------------------
// K1: fused encode transform
// ------------------------------------------------------------------------------------------------
namespace k1 {
constexpr int TW = 128;               // luminance tile, pixels
constexpr int TH = 64;
constexpr int RW = TW + 3;            // staged region: columns/rows -2 .. +T (pyrDown halo)
constexpr int RH = TH + 3;
constexpr int RGB_PITCH = 396;        // RW * 3 = 393, rounded up to a multiple of 4
constexpr int C_PITCH = 132;          // chroma staging pitch
constexpr int CW = TW / 2;            // chroma tile
constexpr int CH = TH / 2;
constexpr int NY_BLOCKS = (TW / 8) * (TH / 8);       // 128
constexpr int NC_BLOCKS = (CW / 8) * (CH / 8);       // 32
constexpr int THREADS = NY_BLOCKS + 2 * NC_BLOCKS;   // 192: one thread per 8x8 block
constexpr float MAGIC = 12582912.0f;                 // 1.5 * 2^23: float add rounds to integer, RN-even

struct Smem {
    union alignas(16) {
        uint8_t rgb[RH * RGB_PITCH];                 // stage 0/1
        uint16_t hpass[2][RH][CW];                   // stage 2 (rgb is dead by then)
    };
    alignas(16) uint8_t y[TH][TW];
    alignas(16) uint8_t cr[RH][C_PITCH];
    alignas(16) uint8_t cb[RH][C_PITCH];
    alignas(16) uint8_t crd[CH][CW];
    alignas(16) uint8_t cbd[CH][CW];
};

// 8x8 block in registers -> quantised zigzag int16, with the near-tie mask.
// KIND: 0 luminance table, 1 chroma table.  v[8*r + c] holds x - 128 (0 where padded).
// umax/vmax: coefficients with u >= umax or v >= vmax lie outside the unpadded plane -> 0.
template <int KIND>
__device__ __forceinline__ void transform_block(float (&v)[64], float abs_sum, int umax, int vmax,
                                                int16_t* __restrict__ dst, uint32_t block_index,
                                                hic_tie_record* __restrict__ ties, uint32_t tie_capacity,
                                                uint32_t* __restrict__ stats) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
        aan_forward8(v[8 * r + 0], v[8 * r + 1], v[8 * r + 2], v[8 * r + 3], v[8 * r + 4], v[8 * r + 5],
                     v[8 * r + 6], v[8 * r + 7]);
#pragma unroll
    for (int c = 0; c < 8; ++c)
        aan_forward8(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);

    constexpr uint8_t zz[64] = HIC_ZIGZAG8;
    const float band = (float)(HIC_TIE_KAPPA * 4.0 / 16777216.0) * abs_sum;     // in units of C = q * v
    int bits[64];
    float margin = 1e30f;
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        const int nat = zz[k];
        const float t = fmaf(v[nat], c_tab.rq[KIND][nat], MAGIC);
        bits[k] = __float_as_int(t);
        if (k != 0) {     // DC = 4*sum(x)/q is exact in float32 and cannot land in a wrong tie
            const float d = fmaf(v[nat], c_tab.rq[KIND][nat], MAGIC - t);
            margin = fminf(margin, (0.5f - fabsf(d)) * c_tab.qf[KIND][nat]);
        }
    }
    const bool cropped = (umax < 8) | (vmax < 8);
    if (cropped) {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const int nat = zz[k];
            if ((nat >> 3) >= umax || (nat & 7) >= vmax) bits[k] = __float_as_int(MAGIC);
        }
    }
    int4* out = reinterpret_cast<int4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int4 w;
        w.x = __byte_perm(bits[8 * j + 0], bits[8 * j + 1], 0x5410);
        w.y = __byte_perm(bits[8 * j + 2], bits[8 * j + 3], 0x5410);
        w.z = __byte_perm(bits[8 * j + 4], bits[8 * j + 5], 0x5410);
        w.w = __byte_perm(bits[8 * j + 6], bits[8 * j + 7], 0x5410);
        out[j] = w;
    }
    if (margin <= band) {       // rare: find which scan positions are inside the band
        uint64_t mask = 0;
#pragma unroll
        for (int k = 1; k < 64; ++k) {
            const int nat = zz[k];
            const float t = fmaf(v[nat], c_tab.rq[KIND][nat], MAGIC);
            const float d = fmaf(v[nat], c_tab.rq[KIND][nat], MAGIC - t);
            const bool inside = (nat >> 3) < umax && (nat & 7) < vmax;
            if (inside && (0.5f - fabsf(d)) * c_tab.qf[KIND][nat] <= band) mask |= (1ull << k);
        }
        if (mask) {
            const uint32_t slot = atomicAdd(&stats[0], 1u);
            if (slot < tie_capacity) {
                hic_tie_record rec;
                rec.block = block_index;
                rec.reserved = 0;
                rec.mask = mask;
                ties[slot] = rec;
            } else {
                atomicAdd(&stats[3], 1u);
            }
        }
    }
}

__global__ void __launch_bounds__(THREADS, 3)
forward_kernel(const uint8_t* __restrict__ rgb, int h, int w, hic_dct_geometry g, int16_t* __restrict__ coef,
               hic_tie_record* __restrict__ ties, uint32_t tie_capacity, uint32_t* __restrict__ stats) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const uint8_t* src = rgb + (size_t)img * h * w * 3;

    // stage 0: stage the RGB region (with BORDER_REFLECT_101 at the image edges) in shared memory
    for (int i = tid; i < RH * RW; i += THREADS) {
        const int ry = i / RW, rx = i - ry * RW;
        const int gy = reflect101(y0 - 2 + ry, h), gx = reflect101(x0 - 2 + rx, w);
        const uint8_t* p = src + ((size_t)gy * w + gx) * 3;
        uint8_t* q = &s.rgb[ry * RGB_PITCH + rx * 3];
        q[0] = p[0];
        q[1] = p[1];
        q[2] = p[2];
    }
    __syncthreads();

    // stage 1: RGB -> Y (tile interior), Cr, Cb (tile + halo)
    for (int i = tid; i < RH * RW; i += THREADS) {
        const int ry = i / RW, rx = i - ry * RW;
        const uint8_t* q = &s.rgb[ry * RGB_PITCH + rx * 3];
        int yy, cr, cb;
        rgb_to_ycrcb(q[0], q[1], q[2], yy, cr, cb);
        s.cr[ry][rx] = (uint8_t)cr;
        s.cb[ry][rx] = (uint8_t)cb;
        const int ty = ry - 2, tx = rx - 2;
        if (ty >= 0 && ty < TH && tx >= 0 && tx < TW) s.y[ty][tx] = (uint8_t)yy;
    }
    __syncthreads();

    // stage 2a: horizontal [1 4 6 4 1] at stride 2 (region column 2*cx + k <-> image column 2*(cx0+cx) - 2 + k)
    for (int i = tid; i < 2 * RH * CW; i += THREADS) {
        const int ch = i / (RH * CW);
        const int rem = i - ch * (RH * CW);
        const int ry = rem / CW, cx = rem - ry * CW;
        const uint8_t* row = ch == 0 ? s.cr[ry] : s.cb[ry];
        const int b = 2 * cx;
        s.hpass[ch][ry][cx] = (uint16_t)(row[b] + 4 * row[b + 1] + 6 * row[b + 2] + 4 * row[b + 3] + row[b + 4]);
    }
    __syncthreads();
    // stage 2b: vertical, (sum + 128) >> 8
    for (int i = tid; i < 2 * CH * CW; i += THREADS) {
        const int ch = i / (CH * CW);
        const int rem = i - ch * (CH * CW);
        const int cy = rem / CW, cx = rem - cy * CW;
        const int b = 2 * cy;
        const int sum = s.hpass[ch][b][cx] + 4 * s.hpass[ch][b + 1][cx] + 6 * s.hpass[ch][b + 2][cx] +
                        4 * s.hpass[ch][b + 3][cx] + s.hpass[ch][b + 4][cx];
        const uint8_t val = (uint8_t)((sum + 128) >> 8);
        if (ch == 0) s.crd[cy][cx] = val; else s.cbd[cy][cx] = val;
    }
    __syncthreads();

    // stage 3: one 8x8 block per thread
    float v[64];
    float abs_sum = 0.f;
    if (tid < NY_BLOCKS) {
        const int by = tid / (TW / 8), bx = tid % (TW / 8);
        const int BY = blockIdx.y * (TH / 8) + by, BX = blockIdx.x * (TW / 8) + bx;
        if (BY >= g.nby_l || BX >= g.nbx_l) return;
        const int rows = min(8, h - 8 * BY), cols = min(8, w - 8 * BX);      // valid pixels
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint2 word = *reinterpret_cast<const uint2*>(&s.y[8 * by + r][8 * bx]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t wv = c < 4 ? word.x : word.y;
                float x = (float)(int)((wv >> (8 * (c & 3))) & 0xFF) - 128.f;
                if (r >= rows || c >= cols) x = 0.f;
                v[8 * r + c] = x;
                abs_sum += fabsf(x);
            }
        }
        const uint32_t block_index = (uint32_t)((int64_t)img * g.blocks_per_image + (int64_t)BY * g.nbx_l + BX);
        transform_block<0>(v, abs_sum, rows, cols, coef + (size_t)block_index * 64, block_index, ties,
                           tie_capacity, stats);
    } else {
        const int plane = (tid - NY_BLOCKS) / NC_BLOCKS;          // 0 = Cr, 1 = Cb
        const int local = (tid - NY_BLOCKS) % NC_BLOCKS;
        const int by = local / (CW / 8), bx = local % (CW / 8);
        const int BY = blockIdx.y * (CH / 8) + by, BX = blockIdx.x * (CW / 8) + bx;
        if (BY >= g.nby_c || BX >= g.nbx_c) return;
        const int rows = min(8, g.hc - 8 * BY), cols = min(8, g.wc - 8 * BX);
        const uint8_t (*pl)[CW] = plane == 0 ? s.crd : s.cbd;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint2 word = *reinterpret_cast<const uint2*>(&pl[8 * by + r][8 * bx]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint32_t wv = c < 4 ? word.x : word.y;
                float x = (float)(int)((wv >> (8 * (c & 3))) & 0xFF) - 128.f;
                if (r >= rows || c >= cols) x = 0.f;
                v[8 * r + c] = x;
                abs_sum += fabsf(x);
            }
        }
        const uint32_t block_index = (uint32_t)((int64_t)img * g.blocks_per_image + g.nb_l + (int64_t)plane * g.nb_c +
                                                (int64_t)BY * g.nbx_c + BX);
        transform_block<1>(v, abs_sum, rows, cols, coef + (size_t)block_index * 64, block_index, ties,
                           tie_capacity, stats);
    }
}

// Float64 re-evaluation of every flagged coefficient with scipy's exact operation order.
__global__ void __launch_bounds__(128)
fixup_kernel(const uint8_t* __restrict__ rgb, int h, int w, hic_dct_geometry g, int16_t* __restrict__ coef,
             const hic_tie_record* __restrict__ ties, uint32_t tie_capacity, uint32_t* __restrict__ stats) {
    const uint32_t n_rec = min(stats[0], tie_capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rec; i += gridDim.x * blockDim.x) {
        const hic_tie_record rec = ties[i];
        const int64_t img = rec.block / g.blocks_per_image;
        int64_t local = rec.block - img * g.blocks_per_image;
        const uint8_t* src = rgb + (size_t)img * h * w * 3;
        int16_t px[64];
        int kind;
        if (local < g.nb_l) {
            kind = 0;
            const int BY = (int)(local / g.nbx_l), BX = (int)(local % g.nbx_l);
            for (int r = 0; r < 8; ++r)
                for (int c = 0; c < 8; ++c) {
                    const int y = 8 * BY + r, x = 8 * BX + c;
                    int val = 0;
                    if (y < h && x < w) {
                        const uint8_t* p = src + ((size_t)y * w + x) * 3;
                        int yy, cr, cb;
                        rgb_to_ycrcb(p[0], p[1], p[2], yy, cr, cb);
                        val = yy - 128;
                    }
                    px[8 * r + c] = (int16_t)val;
                }
        } else {
            kind = 1;
            local -= g.nb_l;
            const int plane = (int)(local / g.nb_c);
            local -= (int64_t)plane * g.nb_c;
            const int BY = (int)(local / g.nbx_c), BX = (int)(local % g.nbx_c);
            for (int r = 0; r < 8; ++r)
                for (int c = 0; c < 8; ++c) {
                    const int cy = 8 * BY + r, cx = 8 * BX + c;
                    int val = 0;
                    if (cy < g.hc && cx < g.wc) {
                        int sum = 0;
                        for (int dy = 0; dy < 5; ++dy) {
                            const int wy = dy == 2 ? 6 : ((dy == 1 || dy == 3) ? 4 : 1);
                            const int y = reflect101(2 * cy - 2 + dy, h);
                            for (int dx = 0; dx < 5; ++dx) {
                                const int wx = dx == 2 ? 6 : ((dx == 1 || dx == 3) ? 4 : 1);
                                const int x = reflect101(2 * cx - 2 + dx, w);
                                const uint8_t* p = src + ((size_t)y * w + x) * 3;
                                int yy, cr, cb;
                                rgb_to_ycrcb(p[0], p[1], p[2], yy, cr, cb);
                                sum += wy * wx * (plane == 0 ? cr : cb);
                            }
                        }
                        val = ((sum + 128) >> 8) - 128;
                    }
                    px[8 * r + c] = (int16_t)val;
                }
        }
        int16_t* blk = coef + (size_t)rec.block * 64;
        uint64_t mask = rec.mask;
        uint32_t evaluated = 0, changed = 0;
        while (mask) {
            const int k = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            const int nat = c_zigzag[k];
            const int32_t exact = exact_quantised_coef(px, nat >> 3, nat & 7, c_tab.qi[kind][nat]);
            ++evaluated;
            if ((int32_t)blk[k] != exact) {
                blk[k] = (int16_t)exact;
                ++changed;
            }
        }
        atomicAdd(&stats[1], evaluated);
        if (changed) atomicAdd(&stats[2], changed);
    }
}
}  // namespace k1

// ------------------------------------------------------------------------------------------------
// K7: dequantise + IDCT + 128 + uint8 cast, one 8x8 block per thread
// ------------------------------------------------------------------------------------------------
namespace k7 {
constexpr float MAGIC = 12582912.0f;

template <int KIND>
__device__ __forceinline__ void inverse_block(const int16_t* __restrict__ src, uint8_t* __restrict__ plane, int ph,
                                              int pw, int BY, int BX, uint32_t block_index,
                                              hic_tie_record* __restrict__ ties, uint32_t tie_capacity,
                                              uint32_t* __restrict__ stats) {
    constexpr uint8_t zz[64] = HIC_ZIGZAG8;
    float v[64];
    float energy = 0.f;          // S = sum |coef * q|
    int ac_bits = 0;
    const int4* in = reinterpret_cast<const int4*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int4 wv = __ldg(in + j);
        const int words[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k0 = 8 * j + 2 * e;
            const int lo = (int)(short)(words[e] & 0xFFFF), hi = words[e] >> 16;
            const float flo = (float)lo, fhi = (float)hi;
            v[zz[k0]] = flo * c_tab.dq[KIND][zz[k0]];
            v[zz[k0 + 1]] = fhi * c_tab.dq[KIND][zz[k0 + 1]];
            energy = fmaf(fabsf(flo), c_tab.qf[KIND][zz[k0]], energy);
            energy = fmaf(fabsf(fhi), c_tab.qf[KIND][zz[k0 + 1]], energy);
            ac_bits |= k0 == 0 ? (words[e] & 0xFFFF0000) : words[e];
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
        aan_inverse8(v[8 * r + 0], v[8 * r + 1], v[8 * r + 2], v[8 * r + 3], v[8 * r + 4], v[8 * r + 5],
                     v[8 * r + 6], v[8 * r + 7]);
#pragma unroll
    for (int c = 0; c < 8; ++c)
        aan_inverse8(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);
    const int rows = min(8, ph - 8 * BY), cols = min(8, pw - 8 * BX);
    // distance of every sample to the nearest integer, where the reference's uint8 truncation steps
    float margin = 1e30f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        v[i] += 128.f;
        const float r = (v[i] + MAGIC) - MAGIC;
        margin = fminf(margin, fabsf(v[i] - r));
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (r >= rows) break;
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint32_t b = (uint32_t)(__float2int_rz(v[8 * r + c])) & 0xFFu;   // truncate, wrap
            if (c < 4) w0 |= b << (8 * c); else w1 |= b << (8 * (c - 4));
        }
        uint8_t* dst = plane + (size_t)(8 * BY + r) * pw + 8 * BX;
        if (cols == 8 && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
            *reinterpret_cast<uint2*>(dst) = make_uint2(w0, w1);
        } else {
            for (int c = 0; c < cols; ++c) dst[c] = (uint8_t)((c < 4 ? w0 >> (8 * c) : w1 >> (8 * (c - 4))) & 0xFF);
        }
    }
    // a DC-only block is exact in both float32 and float64 (no rounding happens at all)
    const float band = (float)(HIC_INV_KAPPA * 4.0 / 16777216.0 / 256.0) * energy + 3.0517578125e-5f;
    if (ac_bits != 0 && margin <= band) {
        uint64_t mask = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const float r = (v[i] + MAGIC) - MAGIC;
            if (fabsf(v[i] - r) <= band && (i >> 3) < rows && (i & 7) < cols) mask |= (1ull << i);
        }
        if (mask) {
            const uint32_t slot = atomicAdd(&stats[0], 1u);
            if (slot < tie_capacity) {
                hic_tie_record rec;
                rec.block = block_index;
                rec.reserved = 0;
                rec.mask = mask;
                ties[slot] = rec;
            } else {
                atomicAdd(&stats[3], 1u);
            }
        }
    }
}

__global__ void __launch_bounds__(128)
inverse_kernel(const int16_t* __restrict__ coef, hic_dct_geometry g, int n, uint8_t* __restrict__ yp,
               uint8_t* __restrict__ crp, uint8_t* __restrict__ cbp, hic_tie_record* __restrict__ ties,
               uint32_t tie_capacity, uint32_t* __restrict__ stats) {
    const int64_t total = (int64_t)n * g.blocks_per_image;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int64_t img = gid / g.blocks_per_image;
    int64_t local = gid - img * g.blocks_per_image;
    const int16_t* src = coef + (size_t)gid * 64;
    if (local < g.nb_l) {
        inverse_block<0>(src, yp + (size_t)img * g.h * g.w, g.h, g.w, (int)(local / g.nbx_l), (int)(local % g.nbx_l),
                         (uint32_t)gid, ties, tie_capacity, stats);
    } else {
        local -= g.nb_l;
        const int plane = (int)(local / g.nb_c);
        local -= (int64_t)plane * g.nb_c;
        uint8_t* base = (plane == 0 ? crp : cbp) + (size_t)img * g.hc * g.wc;
        inverse_block<1>(src, base, g.hc, g.wc, (int)(local / g.nbx_c), (int)(local % g.nbx_c), (uint32_t)gid, ties,
                         tie_capacity, stats);
    }
}

// Float64 re-evaluation of every flagged sample with scipy's exact operation order.
__global__ void __launch_bounds__(128)
inverse_fixup_kernel(const int16_t* __restrict__ coef, hic_dct_geometry g, uint8_t* __restrict__ yp,
                     uint8_t* __restrict__ crp, uint8_t* __restrict__ cbp, const hic_tie_record* __restrict__ ties,
                     uint32_t tie_capacity, uint32_t* __restrict__ stats) {
    const uint32_t n_rec = min(stats[0], tie_capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rec; i += gridDim.x * blockDim.x) {
        const hic_tie_record rec = ties[i];
        const int64_t img = rec.block / g.blocks_per_image;
        int64_t local = rec.block - img * g.blocks_per_image;
        int kind = 0, ph = g.h, pw = g.w, nbx = g.nbx_l;
        uint8_t* plane = yp + (size_t)img * g.h * g.w;
        if (local >= g.nb_l) {
            kind = 1;
            local -= g.nb_l;
            const int p = (int)(local / g.nb_c);
            local -= (int64_t)p * g.nb_c;
            ph = g.hc; pw = g.wc; nbx = g.nbx_c;
            plane = (p == 0 ? crp : cbp) + (size_t)img * g.hc * g.wc;
        }
        const int BY = (int)(local / nbx), BX = (int)(local % nbx);
        const int16_t* blk = coef + (size_t)rec.block * 64;
        int32_t cq[64];
        for (int k = 0; k < 64; ++k) {
            const int nat = c_zigzag[k];
            cq[nat] = (int32_t)blk[k] * c_tab.qi[kind][nat];
        }
        uint64_t mask = rec.mask;
        uint32_t evaluated = 0, changed = 0;
        while (mask) {
            const int s = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            const int y = s >> 3, x = s & 7;
            const uint8_t exact = wrap_u8(exact_decoded_sample(cq, y, x));
            uint8_t* dst = plane + (size_t)(8 * BY + y) * pw + 8 * BX + x;
            ++evaluated;
            if (8 * BY + y < ph && 8 * BX + x < pw && *dst != exact) {
                *dst = exact;
                ++changed;
            }
        }
        atomicAdd(&stats[1], evaluated);
        if (changed) atomicAdd(&stats[2], changed);
    }
}

// K8: cv2.pyrUp of both chroma planes + crop + YCrCb -> RGB.  One thread per chroma sample,
// producing the 2x2 output quad.  pyrUp per axis: even = s[i-1] + 6 s[i] + s[i+1],
// odd = 4 (s[i] + s[i+1]); s[-1] = s[1], s[n] = s[n-1]; (v + 32) >> 6 after both axes.
__global__ void __launch_bounds__(256)
upsample_colour_kernel(const uint8_t* __restrict__ yp, const uint8_t* __restrict__ crp,
                       const uint8_t* __restrict__ cbp, hic_dct_geometry g, int n, uint8_t* __restrict__ rgb) {
    const int64_t per = (int64_t)g.hc * g.wc;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per * n) return;
    const int64_t img = gid / per;
    const int64_t rem = gid - img * per;
    const int cy = (int)(rem / g.wc), cx = (int)(rem % g.wc);
    const int ym = cy > 0 ? cy - 1 : (g.hc > 1 ? 1 : 0), yn = cy + 1 < g.hc ? cy + 1 : g.hc - 1;
    const int xm = cx > 0 ? cx - 1 : (g.wc > 1 ? 1 : 0), xn = cx + 1 < g.wc ? cx + 1 : g.wc - 1;
    int up[2][4];      // [channel][2*dy + dx]
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const uint8_t* p = (ch == 0 ? crp : cbp) + (size_t)img * per;
        const int r0[3] = {p[(size_t)ym * g.wc + xm], p[(size_t)ym * g.wc + cx], p[(size_t)ym * g.wc + xn]};
        const int r1[3] = {p[(size_t)cy * g.wc + xm], p[(size_t)cy * g.wc + cx], p[(size_t)cy * g.wc + xn]};
        const int r2[3] = {p[(size_t)yn * g.wc + xm], p[(size_t)yn * g.wc + cx], p[(size_t)yn * g.wc + xn]};
        // horizontal pass on the three rows: even column, odd column
        const int e0 = r0[0] + 6 * r0[1] + r0[2], o0 = 4 * (r0[1] + r0[2]);
        const int e1 = r1[0] + 6 * r1[1] + r1[2], o1 = 4 * (r1[1] + r1[2]);
        const int e2 = r2[0] + 6 * r2[1] + r2[2], o2 = 4 * (r2[1] + r2[2]);
        up[ch][0] = (e0 + 6 * e1 + e2 + 32) >> 6;
        up[ch][1] = (o0 + 6 * o1 + o2 + 32) >> 6;
        up[ch][2] = (4 * (e1 + e2) + 32) >> 6;
        up[ch][3] = (4 * (o1 + o2) + 32) >> 6;
    }
    const uint8_t* ysrc = yp + (size_t)img * g.h * g.w;
    uint8_t* dst = rgb + (size_t)img * g.out_h * g.out_w * 3;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int oy = 2 * cy + dy, ox = 2 * cx + dx;
            int r, gg, b;
            ycrcb_to_rgb(ysrc[(size_t)oy * g.w + ox], up[0][2 * dy + dx], up[1][2 * dy + dx], r, gg, b);
            uint8_t* o = dst + ((size_t)oy * g.out_w + ox) * 3;
            o[0] = (uint8_t)r;
            o[1] = (uint8_t)gg;
            o[2] = (uint8_t)b;
        }
}
}  // namespace k7

// ------------------------------------------------------------------------------------------------
// layout converters
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
blocks_to_planes_kernel(const int16_t* __restrict__ coef, hic_dct_geometry g, int n, int32_t* __restrict__ lum,
                        int32_t* __restrict__ cr, int32_t* __restrict__ cb) {
    // one thread per coefficient, natural position inside the block varies fastest along x
    const int64_t per = g.blocks_per_image * 64;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per * n) return;
    const int64_t img = gid / per;
    const int64_t rem = gid - img * per;
    int64_t blk = rem >> 6;
    const int k = (int)(rem & 63);
    const int nat = c_zigzag[k];
    const int u = nat >> 3, v = nat & 7;
    int32_t* plane;
    int ph, pw, nbx;
    if (blk < g.nb_l) {
        plane = lum + (size_t)img * g.h * g.w; ph = g.h; pw = g.w; nbx = g.nbx_l;
    } else {
        blk -= g.nb_l;
        const int p = (int)(blk / g.nb_c);
        blk -= (int64_t)p * g.nb_c;
        plane = (p == 0 ? cr : cb) + (size_t)img * g.hc * g.wc; ph = g.hc; pw = g.wc; nbx = g.nbx_c;
    }
    const int y = 8 * (int)(blk / nbx) + u, x = 8 * (int)(blk % nbx) + v;
    if (y < ph && x < pw) plane[(size_t)y * pw + x] = (int32_t)coef[gid];
}

__global__ void __launch_bounds__(256)
planes_to_blocks_kernel(const int32_t* __restrict__ lum, const int32_t* __restrict__ cr,
                        const int32_t* __restrict__ cb, hic_dct_geometry g, int n, int16_t* __restrict__ coef) {
    const int64_t per = g.blocks_per_image * 64;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per * n) return;
    const int64_t img = gid / per;
    const int64_t rem = gid - img * per;
    int64_t blk = rem >> 6;
    const int k = (int)(rem & 63);
    const int nat = c_zigzag[k];
    const int u = nat >> 3, v = nat & 7;
    const int32_t* plane;
    int ph, pw, nbx;
    if (blk < g.nb_l) {
        plane = lum + (size_t)img * g.h * g.w; ph = g.h; pw = g.w; nbx = g.nbx_l;
    } else {
        blk -= g.nb_l;
        const int p = (int)(blk / g.nb_c);
        blk -= (int64_t)p * g.nb_c;
        plane = (p == 0 ? cr : cb) + (size_t)img * g.hc * g.wc; ph = g.hc; pw = g.wc; nbx = g.nbx_c;
    }
    const int y = 8 * (int)(blk / nbx) + u, x = 8 * (int)(blk % nbx) + v;
    coef[gid] = (y < ph && x < pw) ? (int16_t)plane[(size_t)y * pw + x] : (int16_t)0;
}

static inline unsigned grid_for(int64_t items, int threads) { return (unsigned)((items + threads - 1) / threads); }

}  // namespace hic

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int hic_dct_geometry_of(int32_t h, int32_t w, hic_dct_geometry* out) { return hic::geometry_of(h, w, out); }

int hic_dct_forward(const uint8_t* d_rgb, int32_t n, int32_t h, int32_t w, int16_t* d_coef, hic_tie_record* d_ties,
                    uint32_t tie_capacity, uint32_t* d_stats, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_rgb && d_coef && d_ties && d_stats, "NULL device pointer");
    HIC_REQUIRE(n >= 1 && n <= 65535, "batch size must be in 1..65535 (got %d)", n);
    hic_dct_geometry g;
    int rc = geometry_of(h, w, &g);
    if (rc) return rc;
    HIC_REQUIRE((int64_t)n * g.blocks_per_image < (1ll << 32), "batch too large: %lld blocks", (long long)n * g.blocks_per_image);
    rc = ensure_tables();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    HIC_CUDA(cudaMemsetAsync(d_stats, 0, HIC_TIE_STATS * sizeof(uint32_t), st));
    const int ext_x = max(8 * g.nbx_l, 16 * g.nbx_c), ext_y = max(8 * g.nby_l, 16 * g.nby_c);
    dim3 grid(ceil_div(ext_x, k1::TW), ceil_div(ext_y, k1::TH), n);
    static bool attr_set[64] = {false};
    int dev = 0;
    HIC_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        HIC_CUDA(cudaFuncSetAttribute(k1::forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(k1::Smem)));
        if (dev < 64) attr_set[dev] = true;
    }
    HIC_LAUNCH("forward_kernel", st, k1::forward_kernel<<<grid, k1::THREADS, sizeof(k1::Smem), st>>>(d_rgb, h, w, g, d_coef, d_ties, tie_capacity, d_stats));
    HIC_LAUNCH("fixup_kernel", st, k1::fixup_kernel<<<148 * 4, 128, 0, st>>>(d_rgb, h, w, g, d_coef, d_ties, tie_capacity, d_stats));
    return HIC_OK;
}

int hic_blocks_to_planes(const int16_t* d_coef, int32_t n, int32_t h, int32_t w, int32_t* d_lum, int32_t* d_cr,
                         int32_t* d_cb, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_coef && d_lum && d_cr && d_cb, "NULL device pointer");
    HIC_REQUIRE(n >= 1, "batch size must be positive");
    hic_dct_geometry g;
    int rc = geometry_of(h, w, &g);
    if (rc) return rc;
    const int64_t items = (int64_t)n * g.blocks_per_image * 64;
    cudaStream_t st = as_stream(stream);
    HIC_LAUNCH("blocks_to_planes_kernel", st, blocks_to_planes_kernel<<<grid_for(items, 256), 256, 0, st>>>(d_coef, g, n, d_lum, d_cr, d_cb));
    return HIC_OK;
}

int hic_planes_to_blocks(const int32_t* d_lum, const int32_t* d_cr, const int32_t* d_cb, int32_t n, int32_t h,
                         int32_t w, int16_t* d_coef, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_coef && d_lum && d_cr && d_cb, "NULL device pointer");
    HIC_REQUIRE(n >= 1, "batch size must be positive");
    hic_dct_geometry g;
    int rc = geometry_of(h, w, &g);
    if (rc) return rc;
    const int64_t items = (int64_t)n * g.blocks_per_image * 64;
    cudaStream_t st = as_stream(stream);
    HIC_LAUNCH("planes_to_blocks_kernel", st, planes_to_blocks_kernel<<<grid_for(items, 256), 256, 0, st>>>(d_lum, d_cr, d_cb, g, n, d_coef));
    return HIC_OK;
}

int hic_dct_inverse(const int16_t* d_coef, int32_t n, int32_t h, int32_t w, uint8_t* d_y, uint8_t* d_cr,
                    uint8_t* d_cb, uint8_t* d_rgb_out, hic_tie_record* d_ties, uint32_t tie_capacity,
                    uint32_t* d_stats, void* stream) {
    using namespace hic;
    HIC_REQUIRE(d_coef && d_y && d_cr && d_cb && d_rgb_out && d_ties && d_stats, "NULL device pointer");
    HIC_REQUIRE(n >= 1, "batch size must be positive");
    hic_dct_geometry g;
    int rc = geometry_of(h, w, &g);
    if (rc) return rc;
    HIC_REQUIRE((int64_t)n * g.blocks_per_image < (1ll << 32), "batch too large: %lld blocks", (long long)n * g.blocks_per_image);
    rc = ensure_tables();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    HIC_CUDA(cudaMemsetAsync(d_stats, 0, HIC_TIE_STATS * sizeof(uint32_t), st));
    const int64_t blocks = (int64_t)n * g.blocks_per_image;
    HIC_LAUNCH("inverse_kernel", st, k7::inverse_kernel<<<grid_for(blocks, 128), 128, 0, st>>>(d_coef, g, n, d_y, d_cr, d_cb, d_ties, tie_capacity, d_stats));
    HIC_LAUNCH("inverse_fixup_kernel", st, k7::inverse_fixup_kernel<<<148 * 4, 128, 0, st>>>(d_coef, g, d_y, d_cr, d_cb, d_ties, tie_capacity, d_stats));
    const int64_t quads = (int64_t)n * g.hc * g.wc;
    HIC_LAUNCH("upsample_colour_kernel", st, k7::upsample_colour_kernel<<<grid_for(quads, 256), 256, 0, st>>>(d_y, d_cr, d_cb, g, n, d_rgb_out));
    return HIC_OK;
}

}  // extern "C"
