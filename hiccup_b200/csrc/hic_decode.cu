// hic_decode.cu -- entropy decode stage: Huffman decode (D1), run-length expansion (D2), DC prefix
// sum and de-zigzag into blocks (D3).
//
// The `.hic` format carries no restart offsets (codec.py:319-334), so a bit stream can only be
// entered at its first bit.  D1 decodes each of the 9 n symbol streams with a lookup table on the
// first LUT_BITS bits of the window (codes longer than that fall back to a search of the stream's
// long rows); D2 and D3 are tile scans over the decoded symbols.
#include <algorithm>
#include <vector>
#include "hic_core.cuh"
#include "hic_runtime.cuh"

namespace hic {
namespace dec {

constexpr int LUT_BITS = 11;
constexpr int LUT_SIZE = 1 << LUT_BITS;
constexpr int XT = 2048;             // symbols per expand tile
constexpr int XTHREADS = 256;
constexpr int XSPT = XT / XTHREADS;

struct LongRow {
    uint64_t code;      // left aligned in 64 bits
    int32_t sym;
    uint32_t len;
};

struct Geom {
    hic_stream_layout L;
    int xtiles[3];              // expand tiles per channel (capacity = 64 * nb symbols)
    int xtiles_per_image;
    int dtiles[3];              // DC tiles per channel (capacity = nb)
    int dtiles_per_image;
};

__host__ __device__ inline int64_t cs_block_base(const Geom& g, int img, int c) {
    return (int64_t)img * g.L.blocks_per_image + g.L.block_off[c];
}

// lut[ss][prefix] = sym << 8 | len  (len == 0: no code of <= LUT_BITS bits has this prefix)
__global__ void __launch_bounds__(256)
build_lut_kernel(const uint64_t* __restrict__ row_off, const int32_t* __restrict__ row_sym,
                 const uint8_t* __restrict__ row_len, const uint64_t* __restrict__ row_code, int32_t* __restrict__ lut) {
    const int ss = blockIdx.x;
    int32_t* my = lut + (size_t)ss * LUT_SIZE;
    for (int i = threadIdx.x; i < LUT_SIZE; i += blockDim.x) my[i] = 0;
    __syncthreads();
    const uint64_t r0 = row_off[ss], r1 = row_off[ss + 1];
    for (uint64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        const uint32_t len = row_len[r];
        if (len == 0 || len > LUT_BITS) continue;
        const uint32_t base = (uint32_t)(row_code[r] << (LUT_BITS - len));
        const int32_t entry = (row_sym[r] << 8) | (int32_t)len;
        for (uint32_t j = 0; j < (1u << (LUT_BITS - len)); ++j) my[base + j] = entry;
    }
}

// D1: one thread per symbol stream.  Threads are ordered (channel, kind) major, image minor, so the
// lanes of a warp decode streams of similar length.
__global__ void __launch_bounds__(32)
huffman_decode_kernel(Geom g, const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ byte_off,
                      const uint64_t* __restrict__ nbits_arr, const int32_t* __restrict__ lut,
                      const LongRow* __restrict__ long_rows, const uint32_t* __restrict__ long_off,
                      int16_t* __restrict__ dc, int16_t* __restrict__ values, uint8_t* __restrict__ lengths,
                      uint32_t* __restrict__ nsym_out, uint32_t* __restrict__ err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = g.L.n_images;
    if (t >= n * 9) return;
    const int grp = t / n, img = t - grp * n;
    const int c = grp / 3, kind = grp % 3;
    const int ss = (img * 3 + c) * 3 + kind;
    const uint64_t nbits = nbits_arr[ss];
    const int64_t bb = cs_block_base(g, img, c);
    const uint32_t cap = (uint32_t)(kind == HIC_KIND_DC ? g.L.nb[c] : g.L.nb[c] * 64);
    int16_t* out16 = kind == HIC_KIND_DC ? dc + bb : values + bb * 64;
    uint8_t* out8 = lengths + bb * 64;
    const int32_t* my_lut = lut + (size_t)ss * LUT_SIZE;
    const LongRow* lr = long_rows + long_off[ss];
    const uint32_t n_long = long_off[ss + 1] - long_off[ss];

    const uint32_t* words = reinterpret_cast<const uint32_t*>(bytes + byte_off[ss]);    // 4-byte aligned by contract
    uint64_t buf = (uint64_t)__byte_perm(__ldg(words), 0, 0x0123) << 40;                // drop the pad-count byte
    int avail = 24;
    uint32_t next_word = 1;
    uint64_t consumed = 0;
    uint32_t count = 0;
    bool bad = false;
    while (consumed < nbits) {
        if (avail <= 32) {
            buf |= (uint64_t)__byte_perm(__ldg(words + next_word), 0, 0x0123) << (32 - avail);
            ++next_word;
            avail += 32;
        }
        const int32_t e = __ldg(my_lut + (uint32_t)(buf >> (64 - LUT_BITS)));
        int32_t sym;
        uint32_t len = (uint32_t)(e & 0xFF);
        if (len) {
            sym = e >> 8;
        } else {
            // long code: make sure the window holds up to 58 bits, then search the long rows
            if (avail <= 32) {     // (cannot happen right after the refill above, kept for clarity)
                buf |= (uint64_t)__byte_perm(__ldg(words + next_word), 0, 0x0123) << (32 - avail);
                ++next_word;
                avail += 32;
            }
            uint64_t window = buf;
            if (avail < 58) {      // peek one more word without consuming it
                window |= (uint64_t)__byte_perm(__ldg(words + next_word), 0, 0x0123) >> (avail - 32);
            }
            sym = 0;
            for (uint32_t i = 0; i < n_long; ++i) {
                const LongRow r = lr[i];
                if ((window >> (64 - r.len)) == (r.code >> (64 - r.len))) {
                    sym = r.sym;
                    len = r.len;
                    break;
                }
            }
            if (!len) {
                bad = true;
                break;
            }
        }
        if (consumed + len > nbits || count >= cap) {
            bad = true;
            break;
        }
        if (kind == HIC_KIND_LENGTH) out8[count] = (uint8_t)sym;
        else out16[count] = (int16_t)sym;
        ++count;
        consumed += len;
        if (len > (uint32_t)avail) {       // a long code that used the peeked word
            buf = (uint64_t)__byte_perm(__ldg(words + next_word), 0, 0x0123) << 32;
            ++next_word;
            const uint32_t extra = len - avail;
            buf <<= extra;
            avail = 32 - (int)extra;
        } else {
            buf <<= len;
            avail -= (int)len;
        }
    }
    nsym_out[ss] = count;
    if (bad) atomicOr(err, 1u);
}

// ------------------------------------------------------------------------------------------------
// D2: run-length expansion.  position(i) = sum_{j<i} (len_j + 1) + len_i
// ------------------------------------------------------------------------------------------------
struct XRef {
    int img, c, tile;
};
__device__ __forceinline__ XRef locate(const int (&tiles)[3], int per_image, int64_t t) {
    XRef r;
    r.img = (int)(t / per_image);
    int rem = (int)(t - (int64_t)r.img * per_image);
    r.c = 0;
    while (rem >= tiles[r.c]) {
        rem -= tiles[r.c];
        ++r.c;
    }
    r.tile = rem;
    return r;
}

template <int THREADS>
__device__ __forceinline__ int64_t block_excl_sum64(int64_t v, int64_t* smem, int64_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    int64_t base = 0, tot = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
        if (w < warp) base += smem[w];
        tot += smem[w];
    }
    if (total) *total = tot;
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(XTHREADS)
expand_tile_sum_kernel(Geom g, const uint8_t* __restrict__ lengths, const uint32_t* __restrict__ nsym_arr,
                       int64_t* __restrict__ tile_sum) {
    __shared__ int64_t s[XTHREADS / 32];
    const XRef r = locate(g.xtiles, g.xtiles_per_image, blockIdx.x);
    const int ss = (r.img * 3 + r.c) * 3 + HIC_KIND_LENGTH;
    const uint32_t nsym = nsym_arr[ss];
    const uint8_t* len = lengths + cs_block_base(g, r.img, r.c) * 64;
    const uint32_t start = (uint32_t)r.tile * XT + threadIdx.x * XSPT;
    int64_t sum = 0;
#pragma unroll
    for (int j = 0; j < XSPT; ++j)
        if (start + j < nsym) sum += (int64_t)len[start + j] + 1;
    int64_t total;
    block_excl_sum64<XTHREADS>(sum, s, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void stream_scan64_kernel(int n_cs, int tiles0, int tiles1, int tiles2, int per_image,
                                     const int64_t* __restrict__ tile_sum, int64_t* __restrict__ tile_off,
                                     int64_t* __restrict__ stream_total) {
    const int cs = blockIdx.x * blockDim.x + threadIdx.x;
    if (cs >= n_cs) return;
    const int img = cs / 3, c = cs % 3;
    const int tiles[3] = {tiles0, tiles1, tiles2};
    int64_t t0 = (int64_t)img * per_image;
    for (int k = 0; k < c; ++k) t0 += tiles[k];
    int64_t run = 0;
    for (int t = 0; t < tiles[c]; ++t) {
        tile_off[t0 + t] = run;
        run += tile_sum[t0 + t];
    }
    stream_total[cs] = run;
}

__global__ void __launch_bounds__(XTHREADS)
expand_scatter_kernel(Geom g, const int16_t* __restrict__ values, const uint8_t* __restrict__ lengths,
                      const uint32_t* __restrict__ nsym_arr, const int64_t* __restrict__ tile_off,
                      int16_t* __restrict__ coef, uint32_t* __restrict__ err) {
    __shared__ int64_t s[XTHREADS / 32];
    const XRef r = locate(g.xtiles, g.xtiles_per_image, blockIdx.x);
    const int cs = r.img * 3 + r.c;
    const uint32_t nsym = nsym_arr[cs * 3 + HIC_KIND_LENGTH];
    const int64_t bb = cs_block_base(g, r.img, r.c);
    const uint8_t* len = lengths + bb * 64;
    const int16_t* val = values + bb * 64;
    const uint32_t start = (uint32_t)r.tile * XT + threadIdx.x * XSPT;
    int l[XSPT], v[XSPT];
    int64_t sum = 0;
#pragma unroll
    for (int j = 0; j < XSPT; ++j) {
        const bool in = start + j < nsym;
        l[j] = in ? len[start + j] : 0;
        v[j] = in ? val[start + j] : 0;
        if (in) sum += l[j] + 1;
    }
    const int64_t rank = block_excl_sum64<XTHREADS>(sum, s, nullptr);
    int64_t pos = tile_off[blockIdx.x] + rank;
    const int64_t stream_len = g.L.len[r.c];
    int16_t* dst = coef + bb * 64;
#pragma unroll
    for (int j = 0; j < XSPT; ++j) {
        if (start + j >= nsym) break;
        const int64_t p = pos + l[j];
        if (v[j] != 0) {
            if (p >= stream_len) {
                atomicOr(err, 2u);
            } else if (g.L.skip_first) {
                dst[(p / 63) * 64 + (p % 63) + 1] = (int16_t)v[j];
            } else {
                dst[p] = (int16_t)v[j];
            }
        }
        pos = p + 1;
    }
}

// validates each channel stream's expanded length (codec.py:109-111: only a trailing (0,0) may
// leave the array short) and the symbol counts
__global__ void validate_kernel(Geom g, const int16_t* __restrict__ values, const uint8_t* __restrict__ lengths,
                                const uint32_t* __restrict__ nsym_arr, const int64_t* __restrict__ stream_total,
                                uint32_t* __restrict__ err) {
    const int cs = blockIdx.x * blockDim.x + threadIdx.x;
    if (cs >= g.L.n_images * 3) return;
    const int img = cs / 3, c = cs % 3;
    const uint32_t n_val = nsym_arr[cs * 3 + HIC_KIND_VALUE], n_len = nsym_arr[cs * 3 + HIC_KIND_LENGTH];
    if (n_val != n_len || n_len == 0) {
        atomicOr(err, 4u);
        return;
    }
    if (g.L.skip_first && nsym_arr[cs * 3 + HIC_KIND_DC] != (uint32_t)g.L.nb[c]) atomicOr(err, 8u);
    const int64_t bb = cs_block_base(g, img, c);
    const bool trailing = values[bb * 64 + n_len - 1] == 0 && lengths[bb * 64 + n_len - 1] == 0;
    const int64_t total = stream_total[cs];
    if (trailing ? (total > g.L.len[c]) : (total != g.L.len[c])) atomicOr(err, 16u);
}

// ------------------------------------------------------------------------------------------------
// D3: DC prefix sum (utils.invert_differences) written to element 0 of every block
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XTHREADS)
dc_tile_sum_kernel(Geom g, const int16_t* __restrict__ dc, int64_t* __restrict__ tile_sum) {
    __shared__ int64_t s[XTHREADS / 32];
    const XRef r = locate(g.dtiles, g.dtiles_per_image, blockIdx.x);
    const int64_t nb = g.L.nb[r.c];
    const int16_t* src = dc + cs_block_base(g, r.img, r.c);
    const int64_t start = (int64_t)r.tile * XT + threadIdx.x * XSPT;
    int64_t sum = 0;
#pragma unroll
    for (int j = 0; j < XSPT; ++j)
        if (start + j < nb) sum += src[start + j];
    int64_t total;
    block_excl_sum64<XTHREADS>(sum, s, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(XTHREADS)
dc_write_kernel(Geom g, const int16_t* __restrict__ dc, const int64_t* __restrict__ tile_off,
                int16_t* __restrict__ coef) {
    __shared__ int64_t s[XTHREADS / 32];
    const XRef r = locate(g.dtiles, g.dtiles_per_image, blockIdx.x);
    const int64_t nb = g.L.nb[r.c];
    const int64_t bb = cs_block_base(g, r.img, r.c);
    const int16_t* src = dc + bb;
    const int64_t start = (int64_t)r.tile * XT + threadIdx.x * XSPT;
    int d[XSPT];
    int64_t sum = 0;
#pragma unroll
    for (int j = 0; j < XSPT; ++j) {
        d[j] = start + j < nb ? src[start + j] : 0;
        sum += d[j];
    }
    int64_t run = tile_off[blockIdx.x] + block_excl_sum64<XTHREADS>(sum, s, nullptr);
#pragma unroll
    for (int j = 0; j < XSPT; ++j) {
        if (start + j >= nb) break;
        run += d[j];
        coef[(bb + start + j) * 64] = (int16_t)run;
    }
}

}  // namespace dec
}  // namespace hic

using namespace hic;
using namespace hic::dec;

struct hic_decode_plan {
    dec::Geom g;
    int n_cs = 0, n_ss = 0;
    int64_t total_blocks = 0, total_xtiles = 0, total_dtiles = 0;
    int16_t* d_dc = nullptr;
    int16_t* d_values = nullptr;
    uint8_t* d_lengths = nullptr;
    int32_t* d_lut = nullptr;
    LongRow* d_long = nullptr;
    uint64_t long_capacity = 0;
    uint32_t* d_long_off = nullptr;
    uint64_t* d_row_off = nullptr;
    int32_t* d_row_sym = nullptr;
    uint8_t* d_row_len = nullptr;
    uint64_t* d_row_code = nullptr;
    uint64_t row_capacity = 0;
    uint64_t* d_byte_off = nullptr;
    uint64_t* d_nbits = nullptr;
    uint32_t* d_nsym = nullptr;
    uint32_t* d_err = nullptr;
    int64_t* d_tile_sum = nullptr;
    int64_t* d_tile_off = nullptr;
    int64_t* d_stream_total = nullptr;
    bool tables_ready = false;
};

template <typename T>
static cudaError_t dalloc2(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), (count ? count : 1) * sizeof(T));
}

extern "C" {

int hic_decode_plan_destroy(hic_decode_plan* p) {
    if (!p) return HIC_OK;
    void* ptrs[] = {p->d_dc, p->d_values, p->d_lengths, p->d_lut, p->d_long, p->d_long_off, p->d_row_off, p->d_row_sym,
                    p->d_row_len, p->d_row_code, p->d_byte_off, p->d_nbits, p->d_nsym, p->d_err, p->d_tile_sum,
                    p->d_tile_off, p->d_stream_total};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete p;
    return HIC_OK;
}

int hic_decode_plan_create(const hic_stream_layout* L, hic_decode_plan** out) {
    HIC_REQUIRE(out != nullptr && L != nullptr, "NULL argument");
    *out = nullptr;
    HIC_REQUIRE(L->n_images >= 1, "layout has no images");
    hic_decode_plan* p = new hic_decode_plan();
    p->g.L = *L;
    p->g.xtiles_per_image = p->g.dtiles_per_image = 0;
    for (int c = 0; c < 3; ++c) {
        if (L->nb[c] < 1 || L->nb[c] * 64 >= (1ll << 31)) {
            delete p;
            return hic::fail(HIC_ERR_INVALID, "channel stream too long");
        }
        p->g.xtiles[c] = (int)((L->nb[c] * 64 + XT - 1) / XT);
        p->g.dtiles[c] = L->skip_first ? (int)((L->nb[c] + XT - 1) / XT) : 0;
        p->g.xtiles_per_image += p->g.xtiles[c];
        p->g.dtiles_per_image += p->g.dtiles[c];
    }
    p->n_cs = L->n_images * 3;
    p->n_ss = L->n_images * 9;
    p->total_blocks = (int64_t)L->n_images * L->blocks_per_image;
    p->total_xtiles = (int64_t)L->n_images * p->g.xtiles_per_image;
    p->total_dtiles = (int64_t)L->n_images * p->g.dtiles_per_image;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    ok(dalloc2(&p->d_dc, p->total_blocks));
    ok(dalloc2(&p->d_values, p->total_blocks * 64 + 64));
    ok(dalloc2(&p->d_lengths, p->total_blocks * 64 + 64));
    ok(dalloc2(&p->d_lut, (size_t)p->n_ss * LUT_SIZE));
    ok(dalloc2(&p->d_long_off, p->n_ss + 1));
    ok(dalloc2(&p->d_row_off, p->n_ss + 1));
    ok(dalloc2(&p->d_byte_off, p->n_ss));
    ok(dalloc2(&p->d_nbits, p->n_ss));
    ok(dalloc2(&p->d_nsym, p->n_ss));
    ok(dalloc2(&p->d_err, 4));
    ok(dalloc2(&p->d_tile_sum, std::max(p->total_xtiles, p->total_dtiles)));
    ok(dalloc2(&p->d_tile_off, std::max(p->total_xtiles, p->total_dtiles)));
    ok(dalloc2(&p->d_stream_total, p->n_cs));
    if (e != cudaSuccess) {
        hic_decode_plan_destroy(p);
        return hic::fail(HIC_ERR_CUDA, "decode plan allocation failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return HIC_OK;
}

int hic_decode_set_tables(hic_decode_plan* p, const uint32_t* h_rows, const int32_t* h_symbols, const uint8_t* h_lens,
                          const uint64_t* h_codes, void* stream) {
    HIC_REQUIRE(p && h_rows && h_symbols && h_lens && h_codes, "NULL argument");
    cudaStream_t st = as_stream(stream);
    const int nss = p->n_ss;
    std::vector<uint64_t> row_off(nss + 1, 0);
    for (int s = 0; s < nss; ++s) row_off[s + 1] = row_off[s] + h_rows[s];
    const uint64_t total = row_off[nss];
    std::vector<LongRow> longs;
    std::vector<uint32_t> long_off(nss + 1, 0);
    for (int s = 0; s < nss; ++s) {
        long_off[s] = (uint32_t)longs.size();
        const size_t begin = longs.size();
        for (uint64_t r = row_off[s]; r < row_off[s + 1]; ++r) {
            const uint32_t len = h_lens[r];
            HIC_REQUIRE(len >= 1 && len <= 58, "code length %u out of range in stream %d", len, s);
            if (len > (uint32_t)LUT_BITS) longs.push_back(LongRow{h_codes[r] << (64 - len), h_symbols[r], len});
        }
        std::sort(longs.begin() + begin, longs.end(), [](const LongRow& a, const LongRow& b) { return a.len < b.len; });
    }
    long_off[nss] = (uint32_t)longs.size();
    if (total > p->row_capacity) {
        if (p->d_row_sym) cudaFree(p->d_row_sym);
        if (p->d_row_len) cudaFree(p->d_row_len);
        if (p->d_row_code) cudaFree(p->d_row_code);
        p->d_row_sym = nullptr; p->d_row_len = nullptr; p->d_row_code = nullptr;
        p->row_capacity = total + total / 4 + 1024;
        HIC_CUDA(dalloc2(&p->d_row_sym, p->row_capacity));
        HIC_CUDA(dalloc2(&p->d_row_len, p->row_capacity));
        HIC_CUDA(dalloc2(&p->d_row_code, p->row_capacity));
    }
    if (longs.size() > p->long_capacity) {
        if (p->d_long) cudaFree(p->d_long);
        p->d_long = nullptr;
        p->long_capacity = longs.size() + longs.size() / 4 + 1024;
        HIC_CUDA(dalloc2(&p->d_long, p->long_capacity));
    }
    HIC_CUDA(cudaMemcpyAsync(p->d_row_off, row_off.data(), sizeof(uint64_t) * (nss + 1), cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_long_off, long_off.data(), sizeof(uint32_t) * (nss + 1), cudaMemcpyHostToDevice, st));
    if (total) {
        HIC_CUDA(cudaMemcpyAsync(p->d_row_sym, h_symbols, sizeof(int32_t) * total, cudaMemcpyHostToDevice, st));
        HIC_CUDA(cudaMemcpyAsync(p->d_row_len, h_lens, sizeof(uint8_t) * total, cudaMemcpyHostToDevice, st));
        HIC_CUDA(cudaMemcpyAsync(p->d_row_code, h_codes, sizeof(uint64_t) * total, cudaMemcpyHostToDevice, st));
    }
    if (!longs.empty())
        HIC_CUDA(cudaMemcpyAsync(p->d_long, longs.data(), sizeof(LongRow) * longs.size(), cudaMemcpyHostToDevice, st));
    build_lut_kernel<<<nss, 256, 0, st>>>(p->d_row_off, p->d_row_sym, p->d_row_len, p->d_row_code, p->d_lut);
    HIC_CHECK_LAUNCH("build_lut_kernel");
    HIC_CUDA(cudaStreamSynchronize(st));
    p->tables_ready = true;
    return HIC_OK;
}

int hic_decode_run(hic_decode_plan* p, const uint8_t* d_bytes, const uint64_t* h_byte_off, const uint64_t* h_nbits,
                   int16_t* d_coef, void* stream) {
    HIC_REQUIRE(p && d_bytes && h_byte_off && h_nbits && d_coef, "NULL argument");
    HIC_REQUIRE(p->tables_ready, "hic_decode_set_tables has not run");
    HIC_REQUIRE((reinterpret_cast<uintptr_t>(d_bytes) & 3) == 0, "d_bytes must be 4-byte aligned");
    const dec::Geom& g = p->g;
    cudaStream_t st = as_stream(stream);
    const int nss = p->n_ss;
    std::vector<uint64_t> nbits(h_nbits, h_nbits + nss), off(h_byte_off, h_byte_off + nss);
    for (int s = 0; s < nss; ++s)
        HIC_REQUIRE((off[s] & 3) == 0, "stream %d is not 4-byte aligned", s);
    HIC_CUDA(cudaMemcpyAsync(p->d_byte_off, off.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_nbits, nbits.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemsetAsync(p->d_err, 0, 4 * sizeof(uint32_t), st));
    HIC_CUDA(cudaMemsetAsync(d_coef, 0, (size_t)p->total_blocks * 128, st));
    huffman_decode_kernel<<<(nss + 31) / 32, 32, 0, st>>>(g, d_bytes, p->d_byte_off, p->d_nbits, p->d_lut, p->d_long,
                                                         p->d_long_off, p->d_dc, p->d_values, p->d_lengths, p->d_nsym,
                                                         p->d_err);
    HIC_CHECK_LAUNCH("huffman_decode_kernel");
    expand_tile_sum_kernel<<<(unsigned)p->total_xtiles, XTHREADS, 0, st>>>(g, p->d_lengths, p->d_nsym, p->d_tile_sum);
    HIC_CHECK_LAUNCH("expand_tile_sum_kernel");
    stream_scan64_kernel<<<(p->n_cs + 127) / 128, 128, 0, st>>>(p->n_cs, g.xtiles[0], g.xtiles[1], g.xtiles[2],
                                                               g.xtiles_per_image, p->d_tile_sum, p->d_tile_off,
                                                               p->d_stream_total);
    HIC_CHECK_LAUNCH("stream_scan64_kernel");
    expand_scatter_kernel<<<(unsigned)p->total_xtiles, XTHREADS, 0, st>>>(g, p->d_values, p->d_lengths, p->d_nsym,
                                                                         p->d_tile_off, d_coef, p->d_err);
    HIC_CHECK_LAUNCH("expand_scatter_kernel");
    validate_kernel<<<(p->n_cs + 127) / 128, 128, 0, st>>>(g, p->d_values, p->d_lengths, p->d_nsym, p->d_stream_total, p->d_err);
    HIC_CHECK_LAUNCH("validate_kernel");
    if (g.L.skip_first) {
        dc_tile_sum_kernel<<<(unsigned)p->total_dtiles, XTHREADS, 0, st>>>(g, p->d_dc, p->d_tile_sum);
        HIC_CHECK_LAUNCH("dc_tile_sum_kernel");
        stream_scan64_kernel<<<(p->n_cs + 127) / 128, 128, 0, st>>>(p->n_cs, g.dtiles[0], g.dtiles[1], g.dtiles[2],
                                                                   g.dtiles_per_image, p->d_tile_sum, p->d_tile_off,
                                                                   p->d_stream_total);
        HIC_CHECK_LAUNCH("stream_scan64_kernel");
        dc_write_kernel<<<(unsigned)p->total_dtiles, XTHREADS, 0, st>>>(g, p->d_dc, p->d_tile_off, d_coef);
        HIC_CHECK_LAUNCH("dc_write_kernel");
    }
    uint32_t flags[4];
    HIC_CUDA(cudaMemcpyAsync(flags, p->d_err, sizeof(flags), cudaMemcpyDeviceToHost, st));
    HIC_CUDA(cudaStreamSynchronize(st));
    if (flags[0]) return hic::fail(HIC_ERR_CORRUPT, "bit streams did not decode cleanly (flags 0x%x)", flags[0]);
    return HIC_OK;
}

}  // extern "C"
