// hic_entropy.cu -- entropy encode stage: DC differences, run-length symbols, symbol histograms with
// first-occurrence indices (E1), host Huffman construction (E2), bit packing (E3).
//
// Run-length coding as a map + scan.  Reference codec.run_length_coding (codec.py:55-99) walks the
// whole channel with a Python reduce.  Restated per input position p of a channel stream, with
// prev(p) = position of the last non-zero before p (-1 at the start) and last_nz = position of
// the stream's last non-zero:
//     non-zero at p                                   -> symbol ((p - prev - 1) mod 15, value)
//     zero at p, (p - prev) mod 15 == 0, p < last_nz  -> filler (14, 0)
//     after the last position, if last_nz != len - 1  -> one trailing (0, 0)
// which emits exactly the reference's list (l // 15 fillers then (l - 15 (l // 15), value) for a
// run of l zeros; trailing zeros collapse to (0, 0) and their count is dropped).  Every position
// emits at most one symbol, so a tile of positions has a bounded output and the output index is an
// exclusive prefix sum -- no serial dependency besides two small carries per tile.
#include <stdlib.h>
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>
#include "hic_core.cuh"
#include "hic_huffman.cuh"
#include "hic_replay.cuh"
#include "hic_runtime.cuh"

namespace hic {

constexpr int RLE_TB = 256;          // blocks (of 64 elements) per RLE tile = threads per CTA
constexpr int PACK_SPT = 8;          // symbols per thread in the pack kernels
constexpr int PACK_THREADS = 256;
constexpr int PACK_TILE = PACK_SPT * PACK_THREADS;      // 2048 symbols
constexpr int PACK_WORDS = 4096;     // 2048 symbols * 58 bits max + slack, in 32-bit words
constexpr int PACK_SEG_TILES = 8;    // tiles a pack CTA walks one after the other (amortises its fixed costs)
constexpr int PACK_SEG = PACK_SEG_TILES * PACK_TILE;    // 16384 symbols per pack segment = per CTA
constexpr int LEN_BINS = 16;
constexpr uint32_t MAX_CODE_LEN = 58;

struct Segment {     // summary of a run of positions: first/last non-zero and symbols within [first, last]
    int first, last, count;
};
__host__ __device__ inline Segment seg_combine(const Segment& a, const Segment& b) {
    if (b.first < 0) return a;
    if (a.first < 0) return b;
    Segment r;
    r.first = a.first;
    r.last = b.last;
    r.count = a.count + b.count + (b.first - a.last - 1) / 15;
    return r;
}

struct TileCarry {
    int prev_last;          // last non-zero position before the tile (-1: none)
    uint32_t sym_off;       // symbols emitted at positions before the tile
};
struct StreamTotals {
    int last_nz;            // filler limit: last non-zero position of the stream (-1: none), or INT_MAX
                            // when a later row band of the same channel still holds a non-zero
    uint32_t nsym;          // run-length symbols including the trailing (0, 0)
    int first_nz;           // first / last non-zero position inside this stream (-1: none)
    int local_last_nz;
};
// Row-band sharding: a channel stream may be one band of a taller image.  The run-length
// state that crosses the seam is handed in, so that every band emits exactly its slice of the whole
// image's symbol list: `carry_zeros` zero positions since the last non-zero of the previous bands
// (its fillers already belong to those bands), the previous band's last DC value, whether a later band
// still holds a non-zero (then trailing zeros keep emitting fillers and there is no (0, 0) here), and
// whether this band closes the stream (it appends the (0, 0) if the stream ends in zeros).
typedef hic_band_carry BandCarry;
struct CompactEntry {
    int32_t sym;
    uint32_t count, first;
};
struct CompactIndex {
    uint32_t offset, count;
};

struct Geom {               // kernel-side copy of the layout plus derived tile counts
    hic_stream_layout L;
    int tiles[3];
    int tiles_per_image;
    int ptiles[3][3];       // pack segments per (channel, kind)
    int ptiles_per_image;
    int nb_bins;            // value bins
};

__host__ __device__ inline int64_t cs_block_base(const Geom& g, int img, int c) {
    return (int64_t)img * g.L.blocks_per_image + g.L.block_off[c];
}

// ------------------------------------------------------------------------------------------------
// block-wide helpers (256 threads)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Segment warp_reduce_segment(Segment s) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        Segment o;
        o.first = __shfl_down_sync(0xffffffffu, s.first, off);
        o.last = __shfl_down_sync(0xffffffffu, s.last, off);
        o.count = __shfl_down_sync(0xffffffffu, s.count, off);
        if ((threadIdx.x & 31) + off < 32) s = seg_combine(s, o);
    }
    return s;
}

// exclusive scans over the CTA: running maximum (seeded with `seed`) and running sum
template <int THREADS>
__device__ __forceinline__ int block_excl_max(int v, int seed, int* smem /* THREADS/32 ints */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc = max(inc, o);
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    int base = seed;
    for (int w = 0; w < warp; ++w) base = max(base, smem[w]);
    int ex = __shfl_up_sync(0xffffffffu, inc, 1);
    ex = lane == 0 ? base : max(base, ex);
    __syncthreads();
    return ex;
}

template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* smem /* THREADS/32 */, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < THREADS / 32; ++w) {
        if (w < warp) base += smem[w];
        tot += smem[w];
    }
    if (total) *total = tot;
    __syncthreads();
    return base + inc - v;
}

// ------------------------------------------------------------------------------------------------
// E1: run-length symbols
// ------------------------------------------------------------------------------------------------
struct TileRef {
    int img, c, tile;       // tile index inside its channel stream
};
__device__ __forceinline__ TileRef locate_tile(const Geom& g, int64_t t) {
    TileRef r;
    r.img = (int)(t / g.tiles_per_image);
    int rem = (int)(t - (int64_t)r.img * g.tiles_per_image);
    r.c = 0;
    while (rem >= g.tiles[r.c]) {
        rem -= g.tiles[r.c];
        ++r.c;
    }
    r.tile = rem;
    return r;
}

__device__ __forceinline__ void load_block(const int16_t* __restrict__ src, int (&w)[32]) {
    const int4* p = reinterpret_cast<const int4*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int4 v = __ldg(p + j);
        w[4 * j + 0] = v.x;
        w[4 * j + 1] = v.y;
        w[4 * j + 2] = v.z;
        w[4 * j + 3] = v.w;
    }
}
#define HIC_ELEM(w, e) (((e) & 1) ? ((w)[(e) >> 1] >> 16) : (int)(short)((w)[(e) >> 1] & 0xFFFF))

__device__ __forceinline__ void hist_add(uint32_t* __restrict__ hist, uint32_t* __restrict__ first, size_t bin, uint32_t index);

// pass A: per-tile segment summary.  FUSE_DC (whole images, DCT mode): the thread has its block in registers, so the
// DC difference, its store and its histogram entry happen here too -- dc_diff_kernel read every 128-byte block
// again for the one int16 at its head (928 MB of DRAM traffic for 13 MB of DC values on C2).  Row bands cannot
// fuse: a band's first difference needs the DC of the band above, known only after the exchange between the
// two passes.
template <bool SKIP, bool FUSE_DC>
__global__ void __launch_bounds__(RLE_TB)
rle_tile_summary_kernel(const int16_t* __restrict__ coef, Geom g, Segment* __restrict__ tile_seg, int16_t* __restrict__ dc_out,
                        uint32_t* __restrict__ hist, uint32_t* __restrict__ first, uint32_t* __restrict__ err) {
    __shared__ Segment warp_seg[RLE_TB / 32];
    const TileRef tr = locate_tile(g, blockIdx.x);
    const int64_t nb = g.L.nb[tr.c];
    const int64_t b = (int64_t)tr.tile * RLE_TB + threadIdx.x;
    Segment s{-1, -1, 0};
    if (b < nb) {
        int w[32];
        const int64_t block_base = cs_block_base(g, tr.img, tr.c);
        load_block(coef + (block_base + b) * 64, w);
        if (FUSE_DC) {
            const int dc = (int)(short)(w[0] & 0xFFFF);
            const int prev = b > 0 ? (int)__ldg(coef + (block_base + b - 1) * 64) : 0;      // (the neighbouring thread's line)
            const int diff = dc - prev;
            dc_out[block_base + b] = (int16_t)diff;
            const int bin = diff + g.nb_bins / 2;
            if (bin < 0 || bin >= g.nb_bins) atomicOr(err, 1u);
            else hist_add(hist, first, ((size_t)(tr.img * 3 + tr.c) * 3 + HIC_KIND_DC) * g.nb_bins + bin, (uint32_t)b);
        }
        const int len = (int)g.L.len[tr.c];
        const int base = SKIP ? (int)(63 * b) - 1 : (int)(64 * b);       // position of element e is base + e
        int zmod = 0, pend = 0;
#pragma unroll
        for (int e = SKIP ? 1 : 0; e < 64; ++e) {
            const int p = base + e;
            const bool nz = HIC_ELEM(w, e) != 0 && (SKIP || p < len);
            if (nz) {
                if (s.first < 0) s.first = p;
                s.last = p;
                s.count += 1 + pend;
                pend = 0;
                zmod = 0;
            } else if (s.first >= 0) {
                if (++zmod == 15) {
                    zmod = 0;
                    ++pend;
                }
            }
        }
    }
    s = warp_reduce_segment(s);
    if ((threadIdx.x & 31) == 0) warp_seg[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        Segment t = warp_seg[0];
        for (int i = 1; i < RLE_TB / 32; ++i) t = seg_combine(t, warp_seg[i]);
        tile_seg[blockIdx.x] = t;
    }
}

// pass B: one CTA per channel stream scans its tile summaries (the segment monoid), producing the carries and
// stream totals, and writes the trailing (0, 0) symbol (with its histogram contribution).  The virtual non-zero
// a row band inherits from the bands above (`prev0`) is applied to the scanned prefix, not scanned itself.
constexpr int SCAN_TB = 256;
__device__ __forceinline__ Segment seg_shfl_up(const Segment& s, int off) {
    Segment o;
    o.first = __shfl_up_sync(0xffffffffu, s.first, off);
    o.last = __shfl_up_sync(0xffffffffu, s.last, off);
    o.count = __shfl_up_sync(0xffffffffu, s.count, off);
    return o;
}
__global__ void __launch_bounds__(SCAN_TB)
rle_stream_scan_kernel(Geom g, const Segment* __restrict__ tile_seg, const BandCarry* __restrict__ band,
                       TileCarry* __restrict__ carry, StreamTotals* __restrict__ totals,
                       int16_t* __restrict__ values, uint8_t* __restrict__ lengths,
                       uint32_t* __restrict__ hist, uint32_t* __restrict__ first) {
    __shared__ Segment s_warp[SCAN_TB / 32];
    const int cs = blockIdx.x;
    if (cs >= g.L.n_images * 3) return;
    const int img = cs / 3, c = cs % 3;
    int64_t t0 = (int64_t)img * g.tiles_per_image;
    for (int k = 0; k < c; ++k) t0 += g.tiles[k];
    const int per_tile = RLE_TB * (g.L.skip_first ? 63 : 64);
    const BandCarry bc = band ? band[cs] : BandCarry{0, 0, 0, 1};
    const int prev0 = -1 - bc.carry_zeros;       // a virtual non-zero that far back, carrying no symbol
    const int base = bc.carry_zeros / 15;        // fillers inside the carried run were emitted by earlier bands
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Segment none{-1, -1, 0};
    Segment acc = none;                          // every real segment of the tiles before this chunk
    const int n_tiles = g.tiles[c];
    for (int tb = 0; tb < n_tiles; tb += SCAN_TB) {
        const int t = tb + threadIdx.x;
        const Segment mine = t < n_tiles ? tile_seg[t0 + t] : none;
        Segment inc = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const Segment o = seg_shfl_up(inc, off);
            if (lane >= off) inc = seg_combine(o, inc);
        }
        if (lane == 31) s_warp[warp] = inc;
        Segment ex = seg_shfl_up(inc, 1);
        if (lane == 0) ex = none;
        __syncthreads();
        Segment before = acc, chunk = none;
        for (int k = 0; k < SCAN_TB / 32; ++k) {
            if (k < warp) before = seg_combine(before, s_warp[k]);
            chunk = seg_combine(chunk, s_warp[k]);
        }
        const Segment x = seg_combine(before, ex);          // the real segments of all tiles before tile t
        if (t < n_tiles) {
            const int last = x.first >= 0 ? x.last : prev0;
            const int count = x.first >= 0 ? x.count + (x.first - prev0 - 1) / 15 : 0;
            TileCarry tc;
            tc.prev_last = last;
            tc.sym_off = (uint32_t)(count + (t * per_tile - 1 - last) / 15 - base);
            carry[t0 + t] = tc;
        }
        acc = seg_combine(acc, chunk);
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const bool any = acc.first >= 0;
    const int first_nz = any ? acc.first : -1;
    Segment run;
    run.first = prev0;
    run.last = any ? acc.last : prev0;
    run.count = any ? acc.count + (acc.first - prev0 - 1) / 15 : 0;
    const int len = (int)g.L.len[c];
    StreamTotals st;
    st.first_nz = first_nz;
    st.local_last_nz = any ? run.last : -1;
    st.last_nz = bc.more_after ? 0x7FFFFFFF : st.local_last_nz;
    // symbols up to the last local non-zero, then (if a later band continues the run) the fillers of the tail
    uint32_t nsym = (uint32_t)(run.count + (bc.more_after ? (len - 1 - run.last) / 15 : 0) - (any || bc.more_after ? base : 0));
    if (!any && !bc.more_after) nsym = 0;
    const bool ends_in_zeros = (st.local_last_nz != len - 1) || len == 0 || bc.carry_zeros > 0 && !any;
    if (bc.closes_stream && !bc.more_after && ends_in_zeros) {          // trailing zeros -> one (0, 0)
        const int64_t sym_base = cs_block_base(g, img, c) * 64;
        values[sym_base + nsym] = 0;
        lengths[sym_base + nsym] = 0;
        const size_t hv = ((size_t)cs * 3 + HIC_KIND_VALUE) * g.nb_bins + g.nb_bins / 2;
        const size_t hl = ((size_t)cs * 3 + HIC_KIND_LENGTH) * g.nb_bins;
        atomicAdd(&hist[hv], 1u);
        atomicAdd(&hist[hl], 1u);
        atomicMin(&first[hv], nsym);
        atomicMin(&first[hl], nsym);
        ++nsym;
    }
    st.nsym = nsym;
    totals[cs] = st;
}

// first / last non-zero position of every channel stream, from the tile summaries alone (what the host
// needs to work out the seam state of row bands before the emit pass)
__global__ void rle_edges_kernel(Geom g, const Segment* __restrict__ tile_seg, StreamTotals* __restrict__ totals) {
    const int cs = blockIdx.x * blockDim.x + threadIdx.x;
    if (cs >= g.L.n_images * 3) return;
    const int img = cs / 3, c = cs % 3;
    int64_t t0 = (int64_t)img * g.tiles_per_image;
    for (int k = 0; k < c; ++k) t0 += g.tiles[k];
    int first_nz = -1, last_nz = -1;
    for (int t = 0; t < g.tiles[c]; ++t) {
        const Segment s = tile_seg[t0 + t];
        if (s.first >= 0) {
            if (first_nz < 0) first_nz = s.first;
            last_nz = s.last;
        }
    }
    StreamTotals st;
    st.first_nz = first_nz;
    st.local_last_nz = last_nz;
    st.last_nz = last_nz;
    st.nsym = 0;
    totals[cs] = st;
}

__device__ __forceinline__ void hist_add(uint32_t* __restrict__ hist, uint32_t* __restrict__ first, size_t bin,
                                         uint32_t index) {
    atomicAdd(&hist[bin], 1u);
    if (first[bin] > index) atomicMin(&first[bin], index);
}

// DC differences (codec.differential_coding): d[0] = DC[0], d[k] = DC[k] - DC[k-1], with their histogram
// and first-occurrence indices.  A kernel of its own, ahead of the run-length passes: the DC alphabets
// are the largest of an image (luminance ~1700 symbols on C2), so their Huffman construction is the
// longest dependent chain of the encoder -- hic_entropy_build_codes_device starts it as soon as this
// kernel and its compaction are done, side by side with rle_emit.
__global__ void __launch_bounds__(RLE_TB)
dc_diff_kernel(const int16_t* __restrict__ coef, Geom g, const BandCarry* __restrict__ band, int16_t* __restrict__ dc_out,
               uint32_t* __restrict__ hist, uint32_t* __restrict__ first, uint32_t* __restrict__ err) {
    const TileRef tr = locate_tile(g, blockIdx.x);
    const int cs = tr.img * 3 + tr.c;
    const int64_t b = (int64_t)tr.tile * RLE_TB + threadIdx.x;
    if (b >= g.L.nb[tr.c]) return;
    const int64_t block_base = cs_block_base(g, tr.img, tr.c);
    const int dc = (int)__ldg(coef + (block_base + b) * 64);
    const int prev = b > 0 ? (int)__ldg(coef + (block_base + b - 1) * 64) : (band ? band[cs].prev_dc : 0);
    const int diff = dc - prev;
    dc_out[block_base + b] = (int16_t)diff;
    const int bin = diff + g.nb_bins / 2;
    if (bin < 0 || bin >= g.nb_bins) atomicOr(err, 1u);
    else hist_add(hist, first, ((size_t)cs * 3 + HIC_KIND_DC) * g.nb_bins + bin, (uint32_t)b);
}

// pass C: emit run-length symbols and their histograms.
// Symbols of the tile are staged in shared memory and written out coalesced (the tile's output is
// one contiguous run).  Histogram updates are aggregated per warp with match.any and accumulated in
// shared memory for the central value bins and all zero-count bins; one flush per tile.
constexpr int EMIT_CENTRAL = 1024;           // value bins [-512, 512) are privatised in shared memory
constexpr int EMIT_CAP = RLE_TB * 64;        // a position emits at most one symbol

struct EmitSmem {
    int16_t val[EMIT_CAP];
    uint8_t len[EMIT_CAP];
    uint32_t hist_v[EMIT_CENTRAL], first_v[EMIT_CENTRAL];
    uint32_t hist_l[LEN_BINS], first_l[LEN_BINS];
    int smax[RLE_TB / 32];
    uint32_t ssum[RLE_TB / 32];
};

template <bool SKIP>
__device__ __forceinline__ void rle_emit_tile(EmitSmem& sm, const uint32_t tile_id, const int16_t* __restrict__ coef, const Geom& g,
                const TileCarry* __restrict__ carry,
                const StreamTotals* __restrict__ totals, int16_t* __restrict__ values,
                uint8_t* __restrict__ lengths, uint32_t* __restrict__ hist, uint32_t* __restrict__ first,
                uint32_t* __restrict__ err) {
    const TileRef tr = locate_tile(g, tile_id);
    const int cs = tr.img * 3 + tr.c;
    const int64_t nb = g.L.nb[tr.c];
    const int64_t b = (int64_t)tr.tile * RLE_TB + threadIdx.x;
    const int64_t block_base = cs_block_base(g, tr.img, tr.c);
    const int len = (int)g.L.len[tr.c];
    const int last_nz = totals[cs].last_nz;
    const TileCarry tc = carry[tile_id];
    const int half = g.nb_bins / 2;
    const size_t hbase = (size_t)cs * 3 * g.nb_bins;

    for (int i = threadIdx.x; i < EMIT_CENTRAL; i += RLE_TB) {
        sm.hist_v[i] = 0;
        sm.first_v[i] = 0xFFFFFFFFu;
    }
    if (threadIdx.x < LEN_BINS) {
        sm.hist_l[threadIdx.x] = 0;
        sm.first_l[threadIdx.x] = 0xFFFFFFFFu;
    }

    int w[32];
    const bool active = b < nb;
    if (active) load_block(coef + (block_base + b) * 64, w);
    else {
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = 0;
    }
    const int base = SKIP ? (int)(63 * b) - 1 : (int)(64 * b);

    // (the DC differences and their histograms are dc_diff_kernel's: they feed the longest Huffman
    // constructions, which start while this kernel is still running)

    // non-zero mask of the block's run-length positions: bit e = element e holds a non-zero
    const int e0 = SKIP ? 1 : 0;
    const int n_valid = active ? (SKIP ? 64 : max(0, min(64, len - (int)(64 * b)))) : 0;       // elements e0 .. n_valid-1 count
    unsigned long long mask = 0;
    if (active) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const unsigned lo = (w[j] & 0xFFFF) != 0, hi = ((unsigned)w[j] >> 16) != 0;
            mask |= (unsigned long long)(lo | (hi << 1)) << (2 * j);
        }
        if (SKIP) mask &= ~1ull;
        if (n_valid < 64) mask &= (1ull << n_valid) - 1ull;
    }
    const int my_last = mask ? base + 63 - __clzll((long long)mask) : (int)0x80000000;   // none: below every carried position
    const int prev = block_excl_max<RLE_TB>(my_last, tc.prev_last, sm.smax);

    // count pass on the mask alone: a gap of `gap` zeros that follows z0 zeros of the same run holds
    // (z0 + gap) / 15 - z0 / 15 fillers inside this block (a filler sits on every 15th zero of a run)
    const int zeros_before = base + e0 - 1 - prev;       // zeros between prev and my first position
    const bool tail_counts = n_valid > e0 && base + n_valid - 1 < last_nz;      // my trailing zeros are inside a run
    uint32_t cnt = 0;
    {
        unsigned long long m = mask;
        int prev_e = e0 - 1;
        if (m) {                                   // the first non-zero continues a run from the blocks before
            const int e = __ffsll((long long)m) - 1;
            m &= m - 1;
            cnt += 1u + (uint32_t)((zeros_before + (e - prev_e - 1)) / 15 - zeros_before / 15);
            prev_e = e;
            // the other non-zeros: one symbol each, plus fillers only where 15 or more zeros lie between two
            // of them -- looked for with shifts (a run of 15 zero bits strictly between the first and the
            // last non-zero); only such a block walks its gaps one by one
            const int top = 63 - __clzll((long long)mask);
            const unsigned long long z = ~mask & ((1ull << top) - 1ull) & ~((2ull << e) - 1ull);
            const unsigned long long z2 = z & (z >> 1), z4 = z2 & (z2 >> 2), z8 = z4 & (z4 >> 4);
            if ((z8 & (z4 >> 8) & (z2 >> 12) & (z >> 14)) == 0ull) {
                cnt += (uint32_t)__popcll(m);
                prev_e = top;
            } else {
                while (m) {
                    const int e2 = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const int gap = e2 - prev_e - 1;
                    cnt += 1u + (gap >= 15 ? (uint32_t)(gap / 15) : 0u);
                    prev_e = e2;
                }
            }
        }
        const int zt = mask ? 0 : zeros_before;
        if (tail_counts) cnt += (uint32_t)((zt + (n_valid - 1 - prev_e)) / 15 - zt / 15);
    }
    uint32_t tile_total;
    const uint32_t rank = block_excl_sum<RLE_TB>(cnt, sm.ssum, &tile_total);

    // emit pass into the staging area: only the non-zeros do any work
    {
        uint32_t local = rank;
        int z0 = zeros_before, prev_e = e0 - 1;
#pragma unroll
        for (int e = e0; e < 64; ++e) {
            if ((mask >> e) & 1ull) {
                int run = z0 + (e - prev_e - 1);
                if (run >= 15) {                                          // fillers (14, 0): rare
                    for (int f = run / 15 - z0 / 15; f > 0; --f) {
                        sm.val[local] = 0;
                        sm.len[local] = 14;
                        ++local;
                    }
                    run %= 15;
                }
                sm.val[local] = (int16_t)HIC_ELEM(w, e);
                sm.len[local] = (uint8_t)run;
                ++local;
                z0 = 0;
                prev_e = e;
            }
        }
        if (tail_counts) {
            for (int f = (z0 + (n_valid - 1 - prev_e)) / 15 - z0 / 15; f > 0; --f) {
                sm.val[local] = 0;
                sm.len[local] = 14;
                ++local;
            }
        }
    }
    __syncthreads();
    // histograms over the staged symbols: plain shared-memory atomics on the privatised bins (lanes that
    // share a bin serialise inside the atomic unit at ~1 cycle per lane, cheaper than aggregating them
    // first with match.any); a first-occurrence index is only sent when it would lower the bin's
    // -- and, in the same pass, their coalesced write-out (the tile's output is one contiguous run)
    const size_t hv = hbase + (size_t)HIC_KIND_VALUE * g.nb_bins;
    const int64_t sym_base = block_base * 64 + tc.sym_off;
    for (uint32_t i = threadIdx.x; i < tile_total; i += RLE_TB) {
        const int sym_val = sm.val[i], sym_len = sm.len[i];
        values[sym_base + i] = (int16_t)sym_val;
        lengths[sym_base + i] = (uint8_t)sym_len;
        const uint32_t idx = tc.sym_off + i;
        atomicAdd(&sm.hist_l[sym_len], 1u);
        if (sm.first_l[sym_len] > idx) atomicMin(&sm.first_l[sym_len], idx);
        const int central = sym_val + EMIT_CENTRAL / 2;
        if (central >= 0 && central < EMIT_CENTRAL) {
            atomicAdd(&sm.hist_v[central], 1u);
            if (sm.first_v[central] > idx) atomicMin(&sm.first_v[central], idx);
        } else {
            const int bin = sym_val + half;
            if (bin < 0 || bin >= g.nb_bins) {
                atomicOr(err, 1u);
            } else {
                atomicAdd(&hist[hv + bin], 1u);
                if (first[hv + bin] > idx) atomicMin(&first[hv + bin], idx);
            }
        }
    }
    __syncthreads();
    // flush the privatised histograms
    for (int i = threadIdx.x; i < EMIT_CENTRAL; i += RLE_TB) {
        const uint32_t c = sm.hist_v[i];
        if (c) {
            const int bin = i - EMIT_CENTRAL / 2 + half;
            if (bin < 0 || bin >= g.nb_bins) {
                atomicOr(err, 1u);
            } else {
                atomicAdd(&hist[hv + bin], c);
                if (first[hv + bin] > sm.first_v[i]) atomicMin(&first[hv + bin], sm.first_v[i]);
            }
        }
    }
    if (threadIdx.x < LEN_BINS) {
        const uint32_t c = sm.hist_l[threadIdx.x];
        const size_t hl = hbase + (size_t)HIC_KIND_LENGTH * g.nb_bins + threadIdx.x;
        if (c) {
            atomicAdd(&hist[hl], c);
            if (first[hl] > sm.first_l[threadIdx.x]) atomicMin(&first[hl], sm.first_l[threadIdx.x]);
        }
    }
}

// Persistent launch: EMIT_CTAS_PER_SM CTAs per SM walk the tiles.  One CTA slot per SM (shared memory and
// threads) is deliberately left free: the DC Huffman constructions -- a few long-running one-warp CTAs
// with their heaps in shared memory -- run beside this kernel and must find room while it is busy.
constexpr int EMIT_CTAS_PER_SM = 3;

template <bool SKIP>
__global__ void __launch_bounds__(RLE_TB)
rle_emit_kernel(const int16_t* __restrict__ coef, Geom g, const BandCarry* __restrict__ band,
                const TileCarry* __restrict__ carry,
                const StreamTotals* __restrict__ totals, int16_t* __restrict__ dc_out, int16_t* __restrict__ values,
                uint8_t* __restrict__ lengths, uint32_t* __restrict__ hist, uint32_t* __restrict__ first,
                uint32_t* __restrict__ err, uint32_t n_tiles) {
    extern __shared__ __align__(16) uint8_t emit_raw[];
    EmitSmem& sm = *reinterpret_cast<EmitSmem*>(emit_raw);
    for (uint32_t tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
        rle_emit_tile<SKIP>(sm, tile_id, coef, g, carry, totals, values, lengths, hist, first, err);
        __syncthreads();            // the staging area and the privatised histograms are reused
    }
}

// ------------------------------------------------------------------------------------------------
// histogram compaction: one CTA per symbol stream
// ------------------------------------------------------------------------------------------------
// stream selection of the per-stream kernels: everything, the DC streams only, the run-length streams only
enum { SEL_ALL = 0, SEL_DC = 1, SEL_AC = 2 };
__device__ __forceinline__ int selected_stream(int sel, int i) {
    return sel == SEL_ALL ? i : (sel == SEL_DC ? 3 * i : 3 * (i >> 1) + 1 + (i & 1));
}
static inline int selected_count(int sel, int n_cs) { return sel == SEL_ALL ? 3 * n_cs : (sel == SEL_DC ? n_cs : 2 * n_cs); }

// The bins it reads are reset on the way (count 0, first occurrence 0xFFFFFFFF), so the histogram
// arrays are clean again for the next batch without a memset over all of them.
__global__ void __launch_bounds__(256)
compact_kernel(Geom g, int sel, uint32_t* __restrict__ hist, uint32_t* __restrict__ first,
               CompactEntry* __restrict__ entries, CompactIndex* __restrict__ index, uint32_t* __restrict__ cursor) {
    __shared__ uint32_t s_count, s_base, s_pos;
    const int ss = selected_stream(sel, blockIdx.x);
    const int kind = ss % 3;
    const int bins = kind == HIC_KIND_LENGTH ? LEN_BINS : g.nb_bins;
    const size_t off = (size_t)ss * g.nb_bins;
    if (threadIdx.x == 0) {
        s_count = 0;
        s_pos = 0;
    }
    __syncthreads();
    uint32_t mine = 0;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) mine += hist[off + i] != 0;
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) {
        s_base = s_count ? atomicAdd(cursor, s_count) : 0;
        index[ss] = CompactIndex{s_base, s_count};
    }
    __syncthreads();
    const int bias = kind == HIC_KIND_LENGTH ? 0 : g.nb_bins / 2;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) {
        const uint32_t cnt = hist[off + i];
        if (cnt) {
            const uint32_t pos = s_base + atomicAdd(&s_pos, 1u);
            entries[pos] = CompactEntry{i - bias, cnt, first[off + i]};
            hist[off + i] = 0;
            first[off + i] = 0xFFFFFFFFu;
        }
    }
}

// scatter the code rows into the dense per-stream lookup table: lut[ss][bin] = len << 58 | code
__global__ void lut_scatter_kernel(Geom g, const int32_t* __restrict__ row_sym, const uint64_t* __restrict__ row_code,
                                   const uint32_t* __restrict__ row_stream, uint64_t n_rows, uint64_t* __restrict__ lut,
                                   uint8_t* __restrict__ lut_len) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const uint32_t ss = row_stream[i];
    const int bias = (ss % 3) == HIC_KIND_LENGTH ? 0 : g.nb_bins / 2;
    lut[(size_t)ss * g.nb_bins + row_sym[i] + bias] = row_code[i];
    lut_len[(size_t)ss * g.nb_bins + row_sym[i] + bias] = (uint8_t)(row_code[i] >> 58);
}

// ------------------------------------------------------------------------------------------------
// E2 on the device: the same heapq replay as csrc/hic_huffman.cuh, in three kernels.
//   huffman_sort_kernel    one CTA per symbol stream: the stream's compacted histogram entries in order of
//                          first occurrence (= the reference's leaf order) -- ranked through a bitmap of the
//                          first-occurrence positions, or sorted by a bitonic network when the stream is too
//                          long for the bitmap; writes the leaf frequencies and the table's symbol column,
//                          and files the stream in a size tier.
//   huffman_replay_kernel  the serial part.  heapify / heappop / heappush with frequency-only
//                          comparisons is one dependent chain per stream, so its throughput is set by
//                          (streams resident per SM) / (latency of one heap level).  One LANE per
//                          stream, every stream's heap in shared memory at 8 bytes per leaf: a warp
//                          carries as many streams as fit in its shared-memory budget, lanes diverge
//                          freely, and the issue slots are shared instead of one warp per stream.
//                          Heap index i lives in slot i + 1, so the children 2p+1, 2p+2 of p are the
//                          16-byte aligned slot pair (2p+2, 2p+3): one LDS.128 per level.
//   huffman_codes_kernel   one CTA per stream: every leaf walks its parent chain to read off its code
//                          (left = first popped = '1'); rows, code lookup table and stream totals.
// Scratch: leaf frequencies reuse the first-occurrence array and parent links (uint16, bit 15 = "I am
// my parent's left child") reuse the histogram array; both are dead once the entries are compacted.
// ------------------------------------------------------------------------------------------------
constexpr int N_TIERS = 21;
__constant__ int c_tier_bound[N_TIERS] = {16, 32, 64, 128, 192, 256, 384, 512, 640, 768, 896,
                                          1024, 1280, 1536, 1792, 2048, 2560, 3072, 4096, 6144, 8192};
static const int h_tier_bound[N_TIERS] = {16, 32, 64, 128, 192, 256, 384, 512, 640, 768, 896,
                                          1024, 1280, 1536, 1792, 2048, 2560, 3072, 4096, 6144, 8192};
constexpr int SORT_THREADS = 256;
constexpr int REPLAY_SMEM_BUDGET = 48 * 1024;
constexpr int REPLAY_NARROW_BUDGET = 44 * 1024;

constexpr int SORT_MAP_WORDS = 8192;          // first-occurrence bitmap of the ranking path: 262 144 symbol positions
constexpr int SORT_SMEM = 8 * SORT_MAP_WORDS; // bitmap + word prefixes (the bitonic fallback needs 6 * 8192 of it)

// Leaf order = order of first occurrence.  First occurrences are distinct symbol positions, so when the stream's
// positions fit a bitmap in shared memory the rank of an entry is a population count: set the bit of every entry's
// first occurrence, prefix-sum the words' counts, and entry i goes to slot prefix[word] + popc(bits below) -- O(n +
// positions / 32) with three barriers, where the bitonic network needs log^2 n of them (66 for a 1700-leaf DC
// alphabet: it was 1.5 ms of a C2 step).  Streams with more than 262 144 symbols (single huge images) still sort.
__global__ void __launch_bounds__(SORT_THREADS)
huffman_sort_kernel(Geom g, int sel, int allow_narrow, const CompactEntry* __restrict__ entries,
                    const CompactIndex* __restrict__ index,
                    uint32_t* __restrict__ leaf_freq, int32_t* __restrict__ row_sym, uint32_t* __restrict__ tier_count,
                    uint32_t* __restrict__ tier_list, int n_ss, uint32_t* __restrict__ err) {
    extern __shared__ __align__(16) uint8_t sort_raw[];
    __shared__ unsigned long long s_total;
    __shared__ uint32_t s_kmax, s_scan[SORT_THREADS / 32];
    const int ss = selected_stream(sel, blockIdx.x);
    const CompactIndex ix = index[ss];
    const int n = (int)ix.count;
    if (n == 0) return;
    if (n > 8192) {
        if (threadIdx.x == 0) atomicOr(err, 2u);
        return;
    }
    const CompactEntry* my = entries + ix.offset;
    if (threadIdx.x == 0) {
        s_total = 0;
        s_kmax = 0;
    }
    __syncthreads();
    {
        unsigned long long my_total = 0;
        uint32_t kmax = 0;
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
            my_total += my[i].count;
            kmax = max(kmax, my[i].first);
        }
        if (my_total) atomicAdd(&s_total, my_total);
        atomicMax(&s_kmax, kmax);
    }
    __syncthreads();
    const uint32_t words = (s_kmax >> 5) + 1u;
    if (words <= (uint32_t)SORT_MAP_WORDS) {
        uint32_t* map = reinterpret_cast<uint32_t*>(sort_raw);
        uint32_t* pre = map + SORT_MAP_WORDS;
        for (uint32_t w = threadIdx.x; w < words; w += SORT_THREADS) map[w] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
            const uint32_t f = my[i].first;
            atomicOr(&map[f >> 5], 1u << (f & 31));
        }
        __syncthreads();
        // exclusive prefix of the words' population counts, SORT_THREADS words at a time
        uint32_t running = 0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (uint32_t w0 = 0; w0 < words; w0 += SORT_THREADS) {
            const uint32_t w = w0 + threadIdx.x;
            const uint32_t c = w < words ? (uint32_t)__popc(map[w]) : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += o;
            }
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            uint32_t base = running;
            for (int k = 0; k < SORT_THREADS / 32; ++k) {
                if (k < warp) base += s_scan[k];
                running += s_scan[k];
            }
            if (w < words) pre[w] = base + inc - c;
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
            const CompactEntry e = my[i];
            const uint32_t r = pre[e.first >> 5] + (uint32_t)__popc(map[e.first >> 5] & ((1u << (e.first & 31)) - 1u));
            leaf_freq[ix.offset + r] = e.count;
            row_sym[ix.offset + r] = e.sym;
        }
    } else {
        int P = 1;
        while (P < n) P <<= 1;
        uint32_t* key = reinterpret_cast<uint32_t*>(sort_raw);
        uint16_t* order = reinterpret_cast<uint16_t*>(sort_raw + 4 * P);
        for (int i = threadIdx.x; i < P; i += SORT_THREADS) {
            key[i] = i < n ? my[i].first : 0xFFFFFFFFu;
            order[i] = (uint16_t)(i < n ? i : 0);
        }
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < P; i += SORT_THREADS) {
                    const int l = i ^ j;
                    if (l > i) {
                        const bool up = (i & k) == 0;
                        const uint32_t a = key[i], b = key[l];
                        if ((a > b) == up) {
                            key[i] = b;
                            key[l] = a;
                            const uint16_t t = order[i];
                            order[i] = order[l];
                            order[l] = t;
                        }
                    }
                }
                __syncthreads();
            }
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
            const CompactEntry e = my[order[i]];
            leaf_freq[ix.offset + i] = e.count;
            row_sym[ix.offset + i] = e.sym;
        }
    }
    if (threadIdx.x == 0 && n >= 2) {
        int t = 0;
        while (c_tier_bound[t] < n) ++t;
        // a stream whose symbols number fewer than 2^18 replays on packed one-word heap entries (hic_replay.cuh):
        // tiers N_TIERS .. 2 N_TIERS - 1
        if (allow_narrow && s_total < (unsigned long long)REPLAY_NARROW_TOTAL) t += N_TIERS;
        const uint32_t pos = atomicAdd(&tier_count[t], 1u);
        tier_list[(size_t)t * n_ss + pos] = (uint32_t)ss;
    }
}

__global__ void __launch_bounds__(32)
huffman_replay_kernel(int tier, int G, int stride_slots, int n_ss, const CompactIndex* __restrict__ index,
                      const uint32_t* __restrict__ leaf_freq, const uint32_t* __restrict__ tier_count,
                      const uint32_t* __restrict__ tier_list, uint16_t* __restrict__ parent) {
    extern __shared__ __align__(16) uint8_t replay_raw[];
    const uint32_t count = tier_count[tier];
    const uint32_t first = blockIdx.x * (uint32_t)G;
    if (first >= count) return;
    const int lane = threadIdx.x;
    const int mine = (int)min((uint32_t)G, count - first);
    // this lane's stream
    uint32_t my_off = 0;
    int my_n = 0;
    if (lane < mine) {
        const CompactIndex ix = index[tier_list[(size_t)tier * n_ss + first + lane]];
        my_off = ix.offset;
        my_n = (int)ix.count;
    }
    // all lanes fill the heaps: leaves in first-occurrence order
    for (int k = 0; k < mine; ++k) {
        const uint32_t off = __shfl_sync(0xffffffffu, my_off, k);
        const int n = __shfl_sync(0xffffffffu, my_n, k);
        uint2* slot = reinterpret_cast<uint2*>(replay_raw) + (size_t)k * stride_slots;
        for (int i = lane; i < n; i += 32) slot[i + 1] = make_uint2((uint32_t)i, leaf_freq[off + i]);
    }
    __syncwarp();
    if (lane >= mine) return;
    uint2* slot = reinterpret_cast<uint2*>(replay_raw) + (size_t)lane * stride_slots;
    const uint4* slot_pair = reinterpret_cast<const uint4*>(slot);       // pair k = slots (2k, 2k+1)
    uint16_t* par = parent + 2 * (size_t)my_off;
    const int n = my_n;
    int size = n;
    // heapq._siftdown(heap, startpos, pos) with the item already in a register
    auto siftdown = [&](int startpos, int pos, uint2 newitem) {
        while (pos > startpos) {
            const int parentpos = (pos - 1) >> 1;
            const uint2 p = slot[parentpos + 1];
            if (!(newitem.y < p.y)) break;
            slot[pos + 1] = p;
            pos = parentpos;
        }
        slot[pos + 1] = newitem;
    };
    // heapq._siftup(heap, pos): the smaller child moves up until a leaf is reached, then the item
    // bubbles back up (CPython's order of comparisons; ties go to the right child).  Both child pairs
    // one level further down are requested before the comparison that picks between them, so the
    // shared-memory latency overlaps the decision chain (indices are clamped, not predicated: a pair
    // past the end is loaded from a valid slot and never used).
    const int pair_lim = (n + 1) >> 1;                                // last pair index inside the lane's slots
    auto siftup = [&](int pos, uint2 newitem) {
        const int endpos = size, startpos = pos;
        int childpos = 2 * pos + 1;
        if (childpos < endpos) {
            uint4 c = slot_pair[pos + 1];                             // slots 2 pos + 2, 2 pos + 3
            while (true) {
                const uint4 gl = slot_pair[min(childpos + 1, pair_lim)];       // children of childpos
                const uint4 gr = slot_pair[min(childpos + 2, pair_lim)];       // children of childpos + 1
                const bool take_right = (childpos + 1 < endpos) && !(c.y < c.w);
                slot[pos + 1] = take_right ? make_uint2(c.z, c.w) : make_uint2(c.x, c.y);
                pos = childpos + (take_right ? 1 : 0);
                childpos = 2 * pos + 1;
                if (childpos >= endpos) break;
                c = take_right ? gr : gl;
            }
        }
        siftdown(startpos, pos, newitem);
    };
    auto pop = [&]() {
        const uint2 lastelt = slot[size];          // heap[size - 1]
        --size;
        if (size > 0) {
            const uint2 ret = slot[1];
            siftup(0, lastelt);
            return ret;
        }
        return lastelt;
    };
    for (int i = n / 2 - 1; i >= 0; --i) siftup(i, slot[i + 1]);     // heapq.heapify
    int next = n;
    while (size > 1) {
        const uint2 l = pop();
        const uint2 r = pop();
        par[l.x] = (uint16_t)(next | 0x8000);
        par[r.x] = (uint16_t)next;
        ++size;
        siftdown(0, size - 1, make_uint2((uint32_t)next, l.y + r.y));      // heapq.heappush
        ++next;
    }
}

// The same replay on packed one-word heap entries (hic_replay.cuh) for the streams filed in the narrow
// tiers: half the shared memory per stream, and (MODE 2) two heap levels per shared-memory round trip.
template <int MODE>
__global__ void __launch_bounds__(32)
huffman_replay_narrow_kernel(int tier, int G, int stride_slots, int n_ss, const CompactIndex* __restrict__ index,
                             const uint32_t* __restrict__ leaf_freq, const uint32_t* __restrict__ tier_count,
                             const uint32_t* __restrict__ tier_list, uint16_t* __restrict__ parent) {
    extern __shared__ __align__(16) uint8_t replay_raw[];
    const uint32_t count = tier_count[N_TIERS + tier];
    const uint32_t first = blockIdx.x * (uint32_t)G;
    if (first >= count) return;
    const int lane = threadIdx.x;
    const int mine = (int)min((uint32_t)G, count - first);
    uint32_t my_off = 0;
    int my_n = 0;
    if (lane < mine) {
        const CompactIndex ix = index[tier_list[(size_t)(N_TIERS + tier) * n_ss + first + lane]];
        my_off = ix.offset;
        my_n = (int)ix.count;
    }
    uint32_t* base = reinterpret_cast<uint32_t*>(replay_raw);
    for (int k = 0; k < mine; ++k) {          // all lanes fill the heaps: leaves in first-occurrence order
        const uint32_t off = __shfl_sync(0xffffffffu, my_off, k);
        const int n = __shfl_sync(0xffffffffu, my_n, k);
        uint32_t* slot = base + (size_t)k * stride_slots;
        for (int i = lane; i < n; i += 32) slot[i + 1] = (leaf_freq[off + i] << REPLAY_ID_BITS) | (uint32_t)i;
    }
    __syncwarp();
    if (lane >= mine) return;
    replay_narrow<MODE>(base + (size_t)lane * stride_slots, my_n, stride_slots, parent + 2 * (size_t)my_off);
}

__global__ void __launch_bounds__(128)
huffman_codes_kernel(Geom g, const CompactIndex* __restrict__ index, const uint32_t* __restrict__ leaf_freq,
                     const uint16_t* __restrict__ parent, const int32_t* __restrict__ row_sym,
                     uint64_t* __restrict__ row_code, uint64_t* __restrict__ lut, uint8_t* __restrict__ lut_len,
                     uint32_t* __restrict__ ss_nsym, uint64_t* __restrict__ ss_nbits, uint32_t* __restrict__ err) {
    __shared__ unsigned long long s_bits;
    __shared__ uint32_t s_nsym;
    const int ss = blockIdx.x;
    const CompactIndex ix = index[ss];
    const int n = (int)ix.count;
    if (n == 0 || n > 8192) {
        if (threadIdx.x == 0) {
            ss_nsym[ss] = 0;
            ss_nbits[ss] = 0;
        }
        return;
    }
    if (threadIdx.x == 0) {
        s_bits = 0;
        s_nsym = 0;
    }
    __syncthreads();
    const int bias = (ss % 3) == HIC_KIND_LENGTH ? 0 : g.nb_bins / 2;
    const uint16_t* par = parent + 2 * (size_t)ix.offset;
    const int root = 2 * n - 2;                    // the last node created (a single leaf has no tree)
    unsigned long long bits_sum = 0;
    uint32_t sym_sum = 0;
    for (int i = threadIdx.x; i < n; i += 128) {
        uint64_t code = 0;
        uint32_t len = 0;
        if (n == 1) {
            code = 1;                              // huffman.py:66-67,181-182: the lone leaf hangs on the left
            len = 1;
        } else {
            int node = i;
            while (node != root && len < 64) {
                const uint32_t p = par[node];
                code |= (uint64_t)(p >> 15) << len;
                ++len;
                node = (int)(p & 0x7FFF);
            }
        }
        if (len > MAX_CODE_LEN) {
            atomicOr(err, 4u);
            len = MAX_CODE_LEN;
            code &= (1ull << MAX_CODE_LEN) - 1;
        }
        const uint64_t packed = ((uint64_t)len << 58) | code;
        const uint32_t f = leaf_freq[ix.offset + i];
        row_code[ix.offset + i] = packed;
        lut[(size_t)ss * g.nb_bins + row_sym[ix.offset + i] + bias] = packed;
        lut_len[(size_t)ss * g.nb_bins + row_sym[ix.offset + i] + bias] = (uint8_t)len;
        bits_sum += (unsigned long long)f * len;
        sym_sum += f;
    }
    atomicAdd(&s_bits, bits_sum);
    atomicAdd(&s_nsym, sym_sum);
    __syncthreads();
    if (threadIdx.x == 0) {
        ss_nsym[ss] = s_nsym;
        ss_nbits[ss] = s_bits;
    }
}

// byte layout of the framed payloads: exclusive scan over the streams (single CTA)
__global__ void __launch_bounds__(1024)
payload_layout_kernel(int n_ss, const uint32_t* __restrict__ ss_nsym, const uint64_t* __restrict__ ss_nbits,
                      uint64_t* __restrict__ byte_off, uint64_t* __restrict__ byte_len, unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_ss; base += 1024) {
        const int s = base + threadIdx.x;
        unsigned long long len = 0;
        if (s < n_ss && ss_nsym[s]) {
            const unsigned long long nb = ss_nbits[s];
            len = 1 + (nb + (8 - (nb & 7))) / 8;
        }
        const unsigned long long al = (len + 3) & ~3ull;
        unsigned long long inc = al;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        unsigned long long pre = s_carry;
        for (int w2 = 0; w2 < warp; ++w2) pre += s_warp[w2];
        if (s < n_ss) {
            byte_off[s] = len ? pre + inc - al : 0;
            byte_len[s] = len;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[0] = s_carry;
}

// ------------------------------------------------------------------------------------------------
// E3: bit packing
// ------------------------------------------------------------------------------------------------
struct PackRef {
    int ss, tile;
    int64_t sym_base;       // element offset of the stream's symbols in its source array
};
__device__ __forceinline__ PackRef locate_pack_tile(const Geom& g, uint32_t t) {
    PackRef r;
    const int img = (int)(t / (uint32_t)g.ptiles_per_image);
    int rem = (int)(t - (uint32_t)img * (uint32_t)g.ptiles_per_image);
    int c = 0, k = 0;
    while (rem >= g.ptiles[c][k]) {
        rem -= g.ptiles[c][k];
        if (++k == 3) {
            k = 0;
            ++c;
        }
    }
    r.ss = (img * 3 + c) * 3 + k;
    r.tile = rem;
    const int64_t bb = cs_block_base(g, img, c);
    r.sym_base = k == HIC_KIND_DC ? bb : bb * 64;
    return r;
}

// the histogram bins of a thread's PACK_SPT = 8 consecutive symbols (positions pos .. pos + 7; the symbol
// arrays carry 64 elements of slack, so whole vectors may be read past the stream's end -- entries at or
// beyond n_valid are garbage and must not be used as indices)
__device__ __forceinline__ void load_bins(int kind, int half, const int16_t* __restrict__ dc,
                                          const int16_t* __restrict__ values, const uint8_t* __restrict__ lengths,
                                          int64_t pos, uint32_t n_valid, int (&bin)[PACK_SPT]) {
    static_assert(PACK_SPT == 8, "vector widths below");
    if (kind == HIC_KIND_LENGTH) {
        const uint2 v = *reinterpret_cast<const uint2*>(lengths + pos);       // 8-byte aligned
#pragma unroll
        for (int j = 0; j < 8; ++j) bin[j] = (int)(((j < 4 ? v.x : v.y) >> (8 * (j & 3))) & 0xFF);
    } else if (kind == HIC_KIND_DC) {          // DC symbols start at an arbitrary block: scalar loads
#pragma unroll
        for (int j = 0; j < 8; ++j) bin[j] = (uint32_t)j < n_valid ? (int)dc[pos + j] + half : 0;
    } else {
        const uint4 v = *reinterpret_cast<const uint4*>(values + pos);        // 16-byte aligned
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) bin[j] = (int)(short)((wv[j >> 1] >> (16 * (j & 1))) & 0xFFFF) + half;
    }
}

__global__ void __launch_bounds__(PACK_THREADS)
pack_tile_bits_kernel(Geom g, const uint8_t* __restrict__ lut_len, const uint32_t* __restrict__ ss_nsym,
                      const int16_t* __restrict__ dc, const int16_t* __restrict__ values,
                      const uint8_t* __restrict__ lengths, uint32_t* __restrict__ tile_bits) {
    __shared__ uint32_t ssum[PACK_THREADS / 32];
    const PackRef pr = locate_pack_tile(g, blockIdx.x);
    const uint32_t nsym = ss_nsym[pr.ss];
    const uint32_t seg_start = (uint32_t)pr.tile * PACK_SEG;
    if (seg_start >= nsym) {            // nothing in this segment (the grid covers the capacity)
        if (threadIdx.x == 0) tile_bits[blockIdx.x] = 0;
        return;
    }
    const uint32_t seg_end = min(nsym, seg_start + (uint32_t)PACK_SEG);
    uint32_t bits = 0;
    const uint8_t* my = lut_len + (size_t)pr.ss * g.nb_bins;
    const int kind = pr.ss % 3, half = g.nb_bins / 2;
    // code lengths only: a byte per symbol from the length table
    for (uint32_t start = seg_start + threadIdx.x * PACK_SPT; start < seg_end; start += PACK_TILE) {
        const uint32_t n_valid = min((uint32_t)PACK_SPT, seg_end - start);
        int bin[PACK_SPT];
        load_bins(kind, half, dc, values, lengths, pr.sym_base + start, n_valid, bin);
        if (n_valid == PACK_SPT) {
#pragma unroll
            for (int j = 0; j < PACK_SPT; ++j) bits += __ldg(my + bin[j]);
        } else {
#pragma unroll
            for (int j = 0; j < PACK_SPT; ++j)
                if ((uint32_t)j < n_valid) bits += __ldg(my + bin[j]);
        }
    }
    uint32_t total;
    block_excl_sum<PACK_THREADS>(bits, ssum, &total);
    if (threadIdx.x == 0) tile_bits[blockIdx.x] = total;
}

__global__ void pack_stream_scan_kernel(Geom g, const uint32_t* __restrict__ tile_bits, const uint32_t* __restrict__ start_bit,
                                        uint64_t* __restrict__ tile_off) {
    const int ss = blockIdx.x * blockDim.x + threadIdx.x;
    if (ss >= g.L.n_images * 9) return;
    const int img = ss / 9, c = (ss % 9) / 3, k = ss % 3;
    int64_t t0 = (int64_t)img * g.ptiles_per_image;
    for (int cc = 0; cc < 3; ++cc)
        for (int kk = 0; kk < 3; ++kk)
            if (cc < c || (cc == c && kk < k)) t0 += g.ptiles[cc][kk];
    uint64_t run = start_bit[ss];           // 8: the pad-count byte comes first (iohelper.py:41-46); 0..7: a row
                                            // band's bits, pre-shifted to their phase in the stitched stream
    for (int t = 0; t < g.ptiles[c][k]; ++t) {
        tile_off[t0 + t] = run;
        run += tile_bits[t0 + t];
    }
}

__global__ void __launch_bounds__(PACK_THREADS)
pack_emit_kernel(Geom g, const uint64_t* __restrict__ lut, const uint32_t* __restrict__ ss_nsym,
                 const uint64_t* __restrict__ ss_nbits, const uint64_t* __restrict__ ss_byte_off,
                 const uint64_t* __restrict__ tile_off, const uint32_t* __restrict__ start_bit,
                 const int16_t* __restrict__ dc,
                 const int16_t* __restrict__ values, const uint8_t* __restrict__ lengths, uint8_t* __restrict__ out) {
    __shared__ uint32_t buf[PACK_WORDS];
    __shared__ uint32_t ssum[PACK_THREADS / 32];
    const PackRef pr = locate_pack_tile(g, blockIdx.x);
    const uint32_t nsym = ss_nsym[pr.ss];
    const uint32_t seg_start = (uint32_t)pr.tile * PACK_SEG;
    if (seg_start >= nsym) return;
    const uint32_t seg_end = min(nsym, seg_start + (uint32_t)PACK_SEG);
    for (int i = threadIdx.x; i < PACK_WORDS; i += PACK_THREADS) buf[i] = 0;
    const uint64_t* my = lut + (size_t)pr.ss * g.nb_bins;
    const int kind = pr.ss % 3, half = g.nb_bins / 2;
    uint32_t* const dst_base = reinterpret_cast<uint32_t*>(out + ss_byte_off[pr.ss]);
    uint64_t g0 = tile_off[blockIdx.x];                    // bit offset of the segment inside the stream's bytes
    if (pr.tile == 0 && threadIdx.x == 0 && start_bit[pr.ss] == 8) {    // pad count p = 8 - (nbits mod 8) in byte 0
        const uint32_t pad = 8u - (uint32_t)(ss_nbits[pr.ss] & 7);
        atomicOr(dst_base, pad);                           // byte 0 of the little-endian word
    }
    // the segment's tiles one after the other, the bit offset running along
    for (uint32_t tile_start = seg_start; tile_start < seg_end; tile_start += PACK_TILE) {
        uint64_t codes[PACK_SPT];
        uint32_t bits = 0;
        const uint32_t start = tile_start + threadIdx.x * PACK_SPT;
        const uint32_t n_valid = start < seg_end ? min((uint32_t)PACK_SPT, seg_end - start) : 0u;
        {
            int bin[PACK_SPT];
            if (n_valid) load_bins(kind, half, dc, values, lengths, pr.sym_base + start, n_valid, bin);
            if (n_valid == PACK_SPT) {
#pragma unroll
                for (int j = 0; j < PACK_SPT; ++j) codes[j] = __ldg(my + bin[j]);
            } else {
#pragma unroll
                for (int j = 0; j < PACK_SPT; ++j) codes[j] = (uint32_t)j < n_valid ? __ldg(my + bin[j]) : 0ull;
            }
#pragma unroll
            for (int j = 0; j < PACK_SPT; ++j) bits += (uint32_t)(codes[j] >> 58);
        }
        uint32_t total;
        const uint32_t rank = block_excl_sum<PACK_THREADS>(bits, ssum, &total);  // its barriers also order the (re)zeroing of buf
        const uint32_t skew = (uint32_t)(g0 & 31);
        uint32_t wi = (skew + rank) >> 5, fill = (skew + rank) & 31;
        uint32_t l_or = 0;
#pragma unroll
        for (int j = 0; j < PACK_SPT; ++j) l_or |= (uint32_t)(codes[j] >> 58);
        if (bits <= 64u && !(l_or & 32u)) {                                 // every code shorter than 32 bits
            // The usual case, without a branch per symbol (lanes would take it at different symbols, so the
            // warp would pay for it at every one): the thread's codes are concatenated in a 64-bit register,
            // left-aligned, and land in the three words they can touch with one atomic OR each.
            uint32_t ahi = 0, alo = 0;
#pragma unroll
            for (int j = 0; j < PACK_SPT; ++j) {
                const uint32_t l = (uint32_t)(codes[j] >> 58);              // 0 .. 31
                ahi = __funnelshift_l(alo, ahi, l);                         // (ahi:alo) <<= l
                alo = (alo << l) | (uint32_t)codes[j];
            }
            if (bits) {
                // left-align the `bits` valid bits in (ahi:alo)
                const uint32_t up = 64u - bits;                             // 0 .. 63
                if (up >= 32u) {
                    ahi = alo << (up - 32u);
                    alo = 0u;
                } else {
                    ahi = __funnelshift_l(alo, ahi, up);
                    alo = up ? alo << up : alo;
                }
                const uint32_t w0 = ahi >> fill;
                const uint32_t w1 = __funnelshift_r(alo, ahi, fill);        // low word of (ahi:alo) >> fill
                const uint32_t w2 = fill ? alo << (32u - fill) : 0u;
                if (w0) atomicOr(&buf[wi], w0);
                if (w1) atomicOr(&buf[wi + 1], w1);
                if (w2) atomicOr(&buf[wi + 2], w2);
            }
        } else {
            // Long codes: laid into a 64-bit window (hi = the output word being filled, lo = the spill into
            // the next one) aligned with the tile buffer's words, flushed a word at a time.
            uint32_t hi = 0, lo = 0;
            auto put = [&](uint32_t code, uint32_t l) {        // 1 <= l <= 32, code < 2^l
                const uint32_t c = code << (32 - l);
                hi |= c >> fill;
                lo |= __funnelshift_r(0u, c, fill);
                fill += l;
                if (fill >= 32) {
                    atomicOr(&buf[wi], hi);
                    ++wi;
                    hi = lo;
                    lo = 0;
                    fill -= 32;
                }
            };
#pragma unroll
            for (int j = 0; j < PACK_SPT; ++j) {
                const uint32_t l = (uint32_t)(codes[j] >> 58);
                if (l > 32) {                                  // two pieces
                    put((uint32_t)(codes[j] >> 32) & ((1u << 26) - 1u), l - 32);
                    put((uint32_t)codes[j], 32);
                } else if (l) {
                    put((uint32_t)codes[j], l);
                }
            }
            if (fill && hi) atomicOr(&buf[wi], hi);
        }
        __syncthreads();
        const uint32_t n_words = (skew + total + 31) >> 5;
        uint32_t* dst = dst_base + (g0 >> 5);
        for (uint32_t i = threadIdx.x; i < n_words; i += PACK_THREADS) {
            const uint32_t v = __byte_perm(buf[i], 0, 0x0123);      // MSB-first bits -> byte order in memory
            buf[i] = 0;                                            // ready for the next tile
            if (v == 0) continue;
            if (i == 0 || i == n_words - 1) atomicOr(dst + i, v);   // words shared with a neighbouring tile
            else dst[i] = v;
        }
        g0 += total;
        // the next tile's stores into buf come after the barriers of its scan
    }
}

}  // namespace hic

// ------------------------------------------------------------------------------------------------
// plan object and C ABI
// ------------------------------------------------------------------------------------------------
using namespace hic;

struct hic_entropy_plan {
    Geom g;
    int64_t total_tiles = 0, total_ptiles = 0, total_blocks = 0;
    int n_cs = 0, n_ss = 0;
    // device
    Segment* d_tile_seg = nullptr;
    TileCarry* d_carry = nullptr;
    StreamTotals* d_totals = nullptr;
    int16_t* d_dc = nullptr;
    int16_t* d_values = nullptr;
    uint8_t* d_lengths = nullptr;
    uint32_t* d_hist = nullptr;
    uint32_t* d_first = nullptr;
    uint32_t* d_err = nullptr;        // [0] error flags, [1] compaction cursor
    CompactEntry* d_entries = nullptr;
    CompactIndex* d_index = nullptr;
    uint64_t* d_lut = nullptr;
    uint8_t* d_lut_len = nullptr;               // code length per bin (what the bit-count pass needs)
    int32_t* d_row_sym = nullptr;
    uint64_t* d_row_code = nullptr;
    uint32_t* d_row_stream = nullptr;
    uint64_t row_capacity = 0;
    uint32_t* d_ss_nsym = nullptr;
    uint64_t* d_ss_nbits = nullptr;
    uint64_t* d_ss_byte_off = nullptr;
    uint32_t* d_ptile_bits = nullptr;
    uint64_t* d_ptile_off = nullptr;
    uint64_t* d_ss_byte_len = nullptr;
    unsigned long long* d_pay_totals = nullptr; // [0] total payload bytes
    uint32_t* d_start_bit = nullptr;            // per symbol stream: 8 = framed payload, 0..7 = raw band bits at that phase
    BandCarry* d_band = nullptr;                // per channel stream: seam state of row-band sharding (else defaults)
    bool band_mode = false;
    bool start_is_default = true;               // d_start_bit holds 8 everywhere
    uint32_t* d_tier_count = nullptr;           // device Huffman builder: streams per size tier
    uint32_t* d_tier_list = nullptr;            // [tier][n_ss] stream ids
    uint32_t* d_tier_count_dc = nullptr;        // the same for the DC streams, whose construction starts early
    uint32_t* d_tier_list_dc = nullptr;
    uint32_t* d_leaf_freq = nullptr;            // builder scratch: leaf frequencies at the compaction offsets
    uint32_t* d_parent = nullptr;               // builder scratch: two uint16 parent links per entry
    bool hist_clean = false;                    // d_hist / d_first hold their reset values
    bool dc_fused = false;                      // the scan pass of this batch has already produced the DC differences
    cudaEvent_t ev_dc = nullptr;                // DC histograms compacted (recorded by the emit pass)
    bool dc_early = false;                      // the last emit pass recorded ev_dc
    bool prefer_device = false;                 // the last code build ran on the device: start the next DC pass eagerly
    bool dc_launched = false;                   // a DC pass is in flight on aux[0..DC_LANES) and not yet joined
    hic::SmallXfer xfer;                        // page-locked staging of the small transfers (reset after every synchronisation)
    bool device_built = false;                  // codes came from hic_entropy_build_codes_device
    bool host_info_valid = false;               // rows/nsym/nbits/byte_off/byte_len mirror the device
    bool host_tables_valid = false;
    // host results
    std::vector<uint32_t> rows, nsym;
    std::vector<uint64_t> nbits, byte_off, byte_len, row_off, dev_row_start;
    std::vector<int32_t> t_sym;
    std::vector<uint8_t> t_len;
    std::vector<uint64_t> t_code;
    uint64_t total_rows = 0, total_bytes = 0;
    bool codes_ready = false;
    cudaStream_t last_stream = nullptr;         // stream of the last code build (the lazy host mirrors use it)
    // fork/join plumbing of the device Huffman builder
    static constexpr int N_AUX = 7;
    cudaStream_t aux[N_AUX] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_join[N_AUX] = {};
};

static int fill_geom(const hic_stream_layout* L, int value_bins, Geom* g) {
    HIC_REQUIRE(L != nullptr, "layout is NULL");
    HIC_REQUIRE(L->n_images >= 1, "layout has no images");
    HIC_REQUIRE(value_bins >= 64 && value_bins <= 65536 && (value_bins & (value_bins - 1)) == 0,
                "value_bins must be a power of two in 64..65536 (got %d)", value_bins);
    g->L = *L;
    g->nb_bins = value_bins;
    g->tiles_per_image = 0;
    g->ptiles_per_image = 0;
    for (int c = 0; c < 3; ++c) {
        HIC_REQUIRE(L->nb[c] >= 1 && L->nb[c] * 64 < (1ll << 31), "channel stream too long (%lld blocks)", (long long)L->nb[c]);
        HIC_REQUIRE(L->len[c] >= 0 && L->len[c] <= L->nb[c] * (L->skip_first ? 63 : 64), "bad stream length");
        g->tiles[c] = (int)((L->nb[c] + RLE_TB - 1) / RLE_TB);
        g->tiles_per_image += g->tiles[c];
        for (int k = 0; k < 3; ++k) {
            int64_t cap = k == HIC_KIND_DC ? (L->skip_first ? L->nb[c] : 0) : L->nb[c] * 64;
            g->ptiles[c][k] = (int)((cap + PACK_SEG - 1) / PACK_SEG);
            g->ptiles_per_image += g->ptiles[c][k];
        }
    }
    return HIC_OK;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), (count ? count : 1) * sizeof(T));
}

extern "C" {

int hic_layout_dct(int32_t n, int32_t h, int32_t w, hic_stream_layout* out) {
    HIC_REQUIRE(out != nullptr, "layout output is NULL");
    HIC_REQUIRE(n >= 1, "batch size must be positive");
    hic_dct_geometry g;
    int rc = hic_dct_geometry_of(h, w, &g);
    if (rc) return rc;
    out->n_images = n;
    out->skip_first = 1;
    out->blocks_per_image = g.blocks_per_image;
    out->nb[0] = g.nb_l;
    out->nb[1] = out->nb[2] = g.nb_c;
    out->block_off[0] = 0;
    out->block_off[1] = g.nb_l;
    out->block_off[2] = g.nb_l + g.nb_c;
    for (int c = 0; c < 3; ++c) out->len[c] = 63 * out->nb[c];
    return HIC_OK;
}

int hic_layout_flat(int32_t n, int64_t len, hic_stream_layout* out) {
    HIC_REQUIRE(out != nullptr, "layout output is NULL");
    HIC_REQUIRE(n >= 1 && len >= 1, "batch size and length must be positive");
    const int64_t nb = (len + 63) / 64;
    out->n_images = n;
    out->skip_first = 0;
    out->blocks_per_image = 3 * nb;
    for (int c = 0; c < 3; ++c) {
        out->nb[c] = nb;
        out->block_off[c] = c * nb;
        out->len[c] = len;
    }
    return HIC_OK;
}

int hic_entropy_plan_destroy(hic_entropy_plan* p) {
    if (!p) return HIC_OK;
    p->xfer.destroy();
    void* ptrs[] = {p->d_tile_seg, p->d_carry, p->d_totals, p->d_dc, p->d_values, p->d_lengths, p->d_hist, p->d_first,
                    p->d_err, p->d_entries, p->d_index, p->d_lut, p->d_row_sym, p->d_row_code, p->d_row_stream,
                    p->d_ss_nsym, p->d_ss_nbits, p->d_ss_byte_off, p->d_ptile_bits, p->d_ptile_off, p->d_ss_byte_len,
                    p->d_pay_totals, p->d_tier_count, p->d_tier_list, p->d_start_bit, p->d_band, p->d_lut_len,
                    p->d_tier_count_dc, p->d_tier_list_dc, p->d_leaf_freq, p->d_parent};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    for (int a = 0; a < hic_entropy_plan::N_AUX; ++a) {
        if (p->aux[a]) cudaStreamDestroy(p->aux[a]);
        if (p->ev_join[a]) cudaEventDestroy(p->ev_join[a]);
    }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_dc) cudaEventDestroy(p->ev_dc);
    delete p;
    return HIC_OK;
}

int hic_entropy_plan_create(const hic_stream_layout* layout, int32_t value_bins, hic_entropy_plan** out) {
    HIC_REQUIRE(out != nullptr, "plan output is NULL");
    *out = nullptr;
    hic_entropy_plan* p = new hic_entropy_plan();
    int rc = fill_geom(layout, value_bins, &p->g);
    if (rc) {
        delete p;
        return rc;
    }
    const Geom& g = p->g;
    p->n_cs = g.L.n_images * 3;
    p->n_ss = g.L.n_images * 9;
    p->total_tiles = (int64_t)g.L.n_images * g.tiles_per_image;
    p->total_ptiles = (int64_t)g.L.n_images * g.ptiles_per_image;
    p->total_blocks = (int64_t)g.L.n_images * g.L.blocks_per_image;
    const size_t hist_n = (size_t)p->n_ss * g.nb_bins;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    ok(dalloc(&p->d_tile_seg, p->total_tiles));
    ok(dalloc(&p->d_carry, p->total_tiles));
    ok(dalloc(&p->d_totals, p->n_cs));
    ok(dalloc(&p->d_dc, p->total_blocks));
    ok(dalloc(&p->d_values, p->total_blocks * 64 + 64));
    ok(dalloc(&p->d_lengths, p->total_blocks * 64 + 64));
    ok(dalloc(&p->d_hist, hist_n));
    ok(dalloc(&p->d_first, hist_n));
    ok(dalloc(&p->d_err, 4));
    ok(dalloc(&p->d_entries, hist_n));
    ok(dalloc(&p->d_index, p->n_ss));
    ok(dalloc(&p->d_lut, hist_n));
    ok(dalloc(&p->d_lut_len, hist_n));
    ok(dalloc(&p->d_ss_nsym, p->n_ss));
    ok(dalloc(&p->d_ss_nbits, p->n_ss));
    ok(dalloc(&p->d_ss_byte_off, p->n_ss));
    ok(dalloc(&p->d_ptile_bits, p->total_ptiles));
    ok(dalloc(&p->d_ptile_off, p->total_ptiles));
    ok(dalloc(&p->d_ss_byte_len, p->n_ss));
    ok(dalloc(&p->d_pay_totals, 2));
    ok(dalloc(&p->d_start_bit, p->n_ss));
    ok(dalloc(&p->d_band, p->n_cs));
    ok(dalloc(&p->d_tier_count, 2 * N_TIERS));                       // wide tiers, then narrow tiers
    ok(dalloc(&p->d_tier_list, (size_t)2 * N_TIERS * p->n_ss));
    ok(dalloc(&p->d_tier_count_dc, 2 * N_TIERS));
    ok(dalloc(&p->d_tier_list_dc, (size_t)2 * N_TIERS * p->n_ss));
    ok(dalloc(&p->d_leaf_freq, hist_n));
    ok(dalloc(&p->d_parent, hist_n));
    // (default priority: with the highest priority the replay's small shared-memory-heavy CTAs displace the
    // wide kernels of other chunks and the pipelined batch path slows down, 27.6 -> 30.0 ms per C2 batch)
    for (int a = 0; a < hic_entropy_plan::N_AUX; ++a) {
        ok(cudaStreamCreateWithFlags(&p->aux[a], cudaStreamNonBlocking));
        ok(cudaEventCreateWithFlags(&p->ev_join[a], cudaEventDisableTiming));
    }
    ok(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&p->ev_dc, cudaEventDisableTiming));
    if (e == cudaSuccess) {
        ok(cudaMemset(p->d_hist, 0, hist_n * sizeof(uint32_t)));
        ok(cudaMemset(p->d_first, 0xFF, hist_n * sizeof(uint32_t)));
        p->hist_clean = e == cudaSuccess;
    }
    if (e == cudaSuccess) {
        std::vector<uint32_t> start(p->n_ss, 8u);
        ok(cudaMemcpy(p->d_start_bit, start.data(), sizeof(uint32_t) * p->n_ss, cudaMemcpyHostToDevice));
    }
    if (e != cudaSuccess) {
        hic_entropy_plan_destroy(p);
        return hic::fail(HIC_ERR_CUDA, "entropy plan allocation failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return HIC_OK;
}

static int ensure_row_capacity(hic_entropy_plan* p, uint64_t rows);

// framed payloads again after a row-band pack (start bit 8 everywhere)
static int restore_default_start(hic_entropy_plan* p, cudaStream_t st) {
    if (p->start_is_default) return HIC_OK;
    std::vector<uint32_t> start(p->n_ss, 8u);
    HIC_CUDA(cudaMemcpyAsync(p->d_start_bit, start.data(), sizeof(uint32_t) * p->n_ss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaStreamSynchronize(st));
    p->start_is_default = true;
    return HIC_OK;
}

// ---- device Huffman builder plumbing (kernels above; orchestration in hic_entropy_build_codes_device) ----
constexpr int DC_LANES = 4;                     // auxiliary streams that carry the early DC pass

static int builder_prepare(hic_entropy_plan* p) {
    const Geom& g = p->g;
    int rc = ensure_row_capacity(p, (uint64_t)p->n_ss * g.nb_bins);
    if (rc) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    HIC_CUDA(cudaGetDevice(&dev));
    const int max_stride = 8 * (h_tier_bound[N_TIERS - 1] + 2);
    if (dev >= 64 || !attr_set[dev]) {
        HIC_CUDA(cudaFuncSetAttribute(huffman_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_stride));
        HIC_CUDA(cudaFuncSetAttribute(huffman_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM));
        HIC_CUDA(cudaFuncSetAttribute(huffman_replay_narrow_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, REPLAY_SMEM_BUDGET));
        HIC_CUDA(cudaFuncSetAttribute(huffman_replay_narrow_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, REPLAY_SMEM_BUDGET));
        if (dev < 64) attr_set[dev] = true;
    }
    return HIC_OK;
}

// HIC_REPLAY_WIDE (environment): every stream on the two-word entries (the round-1 path, for A/B timing);
// HIC_REPLAY_MODE = 1 | 2: one or two heap levels per step in the packed replay (hic_replay.cuh)
static bool replay_allow_narrow() {
    static const bool on = getenv("HIC_REPLAY_WIDE") == nullptr;
    return on;
}
static int replay_mode() {
    static const int mode = getenv("HIC_REPLAY_MODE") ? atoi(getenv("HIC_REPLAY_MODE")) : 2;
    return mode == 1 ? 1 : 2;
}

// sort + tier filing of a selection of the streams on stream s0
static int builder_sort_pass(hic_entropy_plan* p, int sel, uint32_t* tier_count, uint32_t* tier_list, cudaStream_t s0) {
    const Geom& g = p->g;
    const unsigned grid = (unsigned)selected_count(sel, p->n_cs);
    HIC_CUDA(cudaMemsetAsync(tier_count, 0, 2 * N_TIERS * sizeof(uint32_t), s0));
    const int narrow = replay_allow_narrow() ? 1 : 0;
    HIC_LAUNCH("huffman_sort_kernel", s0, huffman_sort_kernel<<<grid, SORT_THREADS, SORT_SMEM, s0>>>(
        g, sel, narrow, p->d_entries, p->d_index, p->d_leaf_freq, p->d_row_sym, tier_count, tier_list, p->n_ss, p->d_err));
    return HIC_OK;
}

// the replay of every size tier, dealt to the given CUDA streams (longest chains first)
static int builder_replay_pass(hic_entropy_plan* p, const uint32_t* tier_count, const uint32_t* tier_list, cudaStream_t* lanes,
                               int n_lanes) {
    for (int t = N_TIERS - 1; t >= 0; --t) {
        const int stride_slots = h_tier_bound[t] + 2;
        const int stride = 8 * stride_slots;
        const int G = std::max(1, std::min(32, REPLAY_SMEM_BUDGET / stride));
        const unsigned grid = (unsigned)((p->n_ss + G - 1) / G);
        huffman_replay_kernel<<<grid, 32, (size_t)G * stride, lanes[t % n_lanes]>>>(t, G, stride_slots, p->n_ss, p->d_index, p->d_leaf_freq,
                                                                                      tier_count, tier_list,
                                                                                      reinterpret_cast<uint16_t*>(p->d_parent));
        HIC_CHECK_LAUNCH("huffman_replay_kernel");
        if (!replay_allow_narrow()) continue;
        // the narrow tier of the same size: one word per heap entry, and a smaller budget per warp so that
        // five CTAs fit an SM (5 x (44 + 1) KB)
        const int nslots = h_tier_bound[t] + 4;
        const int nstride = 4 * nslots;
        const int NG = std::max(1, std::min(32, REPLAY_NARROW_BUDGET / nstride));
        const unsigned ngrid = (unsigned)((p->n_ss + NG - 1) / NG);
        cudaStream_t lane_st = lanes[(t + n_lanes / 2) % n_lanes];
        if (replay_mode() == 1)
            huffman_replay_narrow_kernel<1><<<ngrid, 32, (size_t)NG * nstride, lane_st>>>(t, NG, nslots, p->n_ss, p->d_index, p->d_leaf_freq, tier_count,
                                                                                          tier_list, reinterpret_cast<uint16_t*>(p->d_parent));
        else
            huffman_replay_narrow_kernel<2><<<ngrid, 32, (size_t)NG * nstride, lane_st>>>(t, NG, nslots, p->n_ss, p->d_index, p->d_leaf_freq, tier_count,
                                                                                          tier_list, reinterpret_cast<uint16_t*>(p->d_parent));
        HIC_CHECK_LAUNCH("huffman_replay_narrow_kernel");
    }
    return HIC_OK;
}

// The DC streams' pass on aux[0..DC_LANES): it hangs on the event the emit pass records after the DC
// compaction, not on the run-length symbols, so it runs beside rle_emit.
static int launch_dc_pass(hic_entropy_plan* p) {
    cudaStream_t* dl = p->aux;
    HIC_CUDA(cudaStreamWaitEvent(dl[0], p->ev_dc, 0));
    int rc = builder_sort_pass(p, SEL_DC, p->d_tier_count_dc, p->d_tier_list_dc, dl[0]);
    if (rc) return rc;
    HIC_CUDA(cudaEventRecord(p->ev_join[0], dl[0]));
    for (int a = 1; a < DC_LANES; ++a) HIC_CUDA(cudaStreamWaitEvent(dl[a], p->ev_join[0], 0));
    rc = builder_replay_pass(p, p->d_tier_count_dc, p->d_tier_list_dc, dl, DC_LANES);
    if (rc) return rc;
    p->dc_launched = true;
    return HIC_OK;
}

// make `st` wait for a DC pass that is still in flight (paths that do not consume it)
static int join_dc_pass(hic_entropy_plan* p, cudaStream_t st) {
    if (!p->dc_launched) return HIC_OK;
    for (int a = 0; a < DC_LANES; ++a) {
        HIC_CUDA(cudaEventRecord(p->ev_join[a], p->aux[a]));
        HIC_CUDA(cudaStreamWaitEvent(st, p->ev_join[a], 0));
    }
    p->dc_launched = false;
    return HIC_OK;
}

static bool serial_build() { return getenv("HIC_ENTROPY_SERIAL") != nullptr; }

static int scan_pass(hic_entropy_plan* p, const int16_t* d_coef, cudaStream_t st, bool fuse_dc = false) {
    const Geom& g = p->g;
    p->codes_ready = false;
    {
        int rc = join_dc_pass(p, st);           // an unconsumed DC pass of the previous batch still reads the entries
        if (rc) return rc;
    }
    const size_t hist_n = (size_t)p->n_ss * g.nb_bins;
    if (!p->hist_clean) {          // normally the compaction of the previous batch has reset every bin it found in use
        HIC_CUDA(cudaMemsetAsync(p->d_hist, 0, hist_n * sizeof(uint32_t), st));
        HIC_CUDA(cudaMemsetAsync(p->d_first, 0xFF, hist_n * sizeof(uint32_t), st));
    }
    p->hist_clean = false;
    HIC_CUDA(cudaMemsetAsync(p->d_err, 0, 4 * sizeof(uint32_t), st));
    const unsigned tiles = (unsigned)p->total_tiles;
    p->dc_fused = fuse_dc && g.L.skip_first;
    if (p->dc_fused)
        HIC_LAUNCH("rle_tile_summary_kernel", st, (rle_tile_summary_kernel<true, true><<<tiles, RLE_TB, 0, st>>>(d_coef, g, p->d_tile_seg, p->d_dc, p->d_hist, p->d_first, p->d_err)));
    else if (g.L.skip_first)
        HIC_LAUNCH("rle_tile_summary_kernel", st, (rle_tile_summary_kernel<true, false><<<tiles, RLE_TB, 0, st>>>(d_coef, g, p->d_tile_seg, p->d_dc, p->d_hist, p->d_first, p->d_err)));
    else
        HIC_LAUNCH("rle_tile_summary_kernel", st, (rle_tile_summary_kernel<false, false><<<tiles, RLE_TB, 0, st>>>(d_coef, g, p->d_tile_seg, p->d_dc, p->d_hist, p->d_first, p->d_err)));
    return HIC_OK;
}

static int emit_pass(hic_entropy_plan* p, const int16_t* d_coef, const BandCarry* d_band, cudaStream_t st) {
    const Geom& g = p->g;
    const unsigned tiles = (unsigned)p->total_tiles;
    p->dc_early = false;
    if (g.L.skip_first) {
        if (!p->dc_fused)
            HIC_LAUNCH("dc_diff_kernel", st, dc_diff_kernel<<<tiles, RLE_TB, 0, st>>>(d_coef, g, d_band, p->d_dc, p->d_hist, p->d_first, p->d_err));
        HIC_LAUNCH("compact_kernel", st, compact_kernel<<<selected_count(SEL_DC, p->n_cs), 256, 0, st>>>(g, SEL_DC, p->d_hist, p->d_first, p->d_entries, p->d_index, p->d_err + 1));
        HIC_CUDA(cudaEventRecord(p->ev_dc, st));
        p->dc_early = true;
        // launched ahead of rle_emit when the codes are going to be built on the device (as they were last
        // time): kernels that are launched first are placed first
        if (p->prefer_device && g.nb_bins <= 8192 && !serial_build()) {
            int rc = builder_prepare(p);
            if (rc) return rc;
            rc = launch_dc_pass(p);
            if (rc) return rc;
        }
    }
    HIC_LAUNCH("rle_stream_scan_kernel", st, rle_stream_scan_kernel<<<p->n_cs, SCAN_TB, 0, st>>>(g, p->d_tile_seg, d_band, p->d_carry, p->d_totals,
                                                                 p->d_values, p->d_lengths, p->d_hist, p->d_first));
    {
        static bool emit_attr[64] = {false};
        int dev = 0;
        HIC_CUDA(cudaGetDevice(&dev));
        if (dev >= 64 || !emit_attr[dev]) {
            HIC_CUDA(cudaFuncSetAttribute(rle_emit_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem)));
            HIC_CUDA(cudaFuncSetAttribute(rle_emit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem)));
            if (dev < 64) emit_attr[dev] = true;
        }
    }
    {
        int dev = 0, sms = 148;
        HIC_CUDA(cudaGetDevice(&dev));
        HIC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const unsigned grid = std::min(tiles, (unsigned)(sms * EMIT_CTAS_PER_SM));
        if (g.L.skip_first)
            HIC_LAUNCH("rle_emit_kernel", st, rle_emit_kernel<true><<<grid, RLE_TB, sizeof(EmitSmem), st>>>(d_coef, g, d_band, p->d_carry, p->d_totals, p->d_dc,
                                                            p->d_values, p->d_lengths, p->d_hist, p->d_first, p->d_err, tiles));
        else
            HIC_LAUNCH("rle_emit_kernel", st, rle_emit_kernel<false><<<grid, RLE_TB, sizeof(EmitSmem), st>>>(d_coef, g, d_band, p->d_carry, p->d_totals, p->d_dc,
                                                             p->d_values, p->d_lengths, p->d_hist, p->d_first, p->d_err, tiles));
    }
    {
        const int sel = g.L.skip_first ? SEL_AC : SEL_ALL;
        HIC_LAUNCH("compact_kernel", st, compact_kernel<<<selected_count(sel, p->n_cs), 256, 0, st>>>(g, sel, p->d_hist, p->d_first, p->d_entries, p->d_index, p->d_err + 1));
    }
    p->hist_clean = true;
    return HIC_OK;
}

int hic_entropy_symbolize(hic_entropy_plan* p, const int16_t* d_coef, void* stream) {
    HIC_REQUIRE(p && d_coef, "NULL argument");
    cudaStream_t st = as_stream(stream);
    p->band_mode = false;
    int rc = scan_pass(p, d_coef, st, true);
    if (rc) return rc;
    return emit_pass(p, d_coef, nullptr, st);
}

int hic_entropy_scan(hic_entropy_plan* p, const int16_t* d_coef, int32_t* h_first_nz, int32_t* h_last_nz, void* stream) {
    HIC_REQUIRE(p && d_coef && h_first_nz && h_last_nz, "NULL argument");
    cudaStream_t st = as_stream(stream);
    int rc = scan_pass(p, d_coef, st);
    if (rc) return rc;
    HIC_LAUNCH("rle_edges_kernel", st, rle_edges_kernel<<<(p->n_cs + 127) / 128, 128, 0, st>>>(p->g, p->d_tile_seg, p->d_totals));
    std::vector<StreamTotals> tot(p->n_cs);
    HIC_CUDA(cudaMemcpyAsync(tot.data(), p->d_totals, sizeof(StreamTotals) * p->n_cs, cudaMemcpyDeviceToHost, st));
    HIC_CUDA(cudaStreamSynchronize(st));
    for (int cs = 0; cs < p->n_cs; ++cs) {
        h_first_nz[cs] = tot[cs].first_nz;
        h_last_nz[cs] = tot[cs].local_last_nz;
    }
    return HIC_OK;
}

int hic_entropy_emit(hic_entropy_plan* p, const int16_t* d_coef, const hic_band_carry* h_band, void* stream) {
    HIC_REQUIRE(p && d_coef, "NULL argument");
    cudaStream_t st = as_stream(stream);
    p->band_mode = h_band != nullptr;
    if (h_band) {
        for (int cs = 0; cs < p->n_cs; ++cs)
            HIC_REQUIRE(h_band[cs].carry_zeros >= 0 && h_band[cs].carry_zeros < (1 << 30), "stream %d: bad carry", cs);
        HIC_CUDA(cudaMemcpyAsync(p->d_band, h_band, sizeof(BandCarry) * p->n_cs, cudaMemcpyHostToDevice, st));
        HIC_CUDA(cudaStreamSynchronize(st));          // the caller's array may go
    }
    return emit_pass(p, d_coef, h_band ? p->d_band : nullptr, st);
}

int hic_entropy_histograms(hic_entropy_plan* p, uint32_t* h_index, int32_t* h_entries, uint64_t capacity, uint64_t* n_entries,
                           uint32_t* h_nsym_rl, void* stream) {
    HIC_REQUIRE(p && h_index && n_entries, "NULL argument");
    cudaStream_t st = as_stream(stream);
    uint32_t flags[4];
    HIC_CUDA(cudaMemcpyAsync(flags, p->d_err, sizeof(flags), cudaMemcpyDeviceToHost, st));
    HIC_CUDA(cudaMemcpyAsync(h_index, p->d_index, sizeof(CompactIndex) * p->n_ss, cudaMemcpyDeviceToHost, st));
    std::vector<StreamTotals> tot(p->n_cs);
    HIC_CUDA(cudaMemcpyAsync(tot.data(), p->d_totals, sizeof(StreamTotals) * p->n_cs, cudaMemcpyDeviceToHost, st));
    HIC_CUDA(cudaStreamSynchronize(st));
    if (flags[0]) return hic::fail(HIC_ERR_INVALID, "a symbol fell outside [-%d, %d): create the plan with more value_bins",
                                   p->g.nb_bins / 2, p->g.nb_bins / 2);
    *n_entries = flags[1];
    if (h_nsym_rl)
        for (int cs = 0; cs < p->n_cs; ++cs) h_nsym_rl[cs] = tot[cs].nsym;
    if (h_entries) {
        if (flags[1] > capacity) return hic::fail(HIC_ERR_CAPACITY, "%u histogram entries, room for %llu", flags[1], (unsigned long long)capacity);
        static_assert(sizeof(CompactEntry) == 12, "entry layout");
        if (flags[1]) {
            HIC_CUDA(cudaMemcpyAsync(h_entries, p->d_entries, sizeof(CompactEntry) * flags[1], cudaMemcpyDeviceToHost, st));
            HIC_CUDA(cudaStreamSynchronize(st));
        }
    }
    return HIC_OK;
}

int hic_entropy_set_codes(hic_entropy_plan* p, const uint32_t* h_index, const int32_t* h_row_sym, const uint64_t* h_row_packed,
                          uint64_t total_rows, const uint32_t* h_nsym, const uint64_t* h_nbits, const uint32_t* h_start_bit,
                          void* stream) {
    HIC_REQUIRE(p && h_index && h_row_sym && h_row_packed && h_nsym && h_nbits, "NULL argument");
    const Geom& g = p->g;
    cudaStream_t st = as_stream(stream);
    p->last_stream = st;
    p->prefer_device = false;
    {
        int rc = join_dc_pass(p, st);
        if (rc) return rc;
    }
    const int nss = p->n_ss;
    std::vector<uint32_t> row_stream(total_rows), start(nss, 8u);
    p->rows.assign(nss, 0); p->nsym.assign(nss, 0); p->nbits.assign(nss, 0);
    p->byte_off.assign(nss, 0); p->byte_len.assign(nss, 0); p->row_off.assign(nss + 1, 0);
    p->dev_row_start.assign(nss, 0);
    uint64_t off = 0;
    for (int s = 0; s < nss; ++s) {
        const uint32_t r0 = h_index[2 * s], cnt = h_index[2 * s + 1];
        HIC_REQUIRE((uint64_t)r0 + cnt <= total_rows, "stream %d: rows [%u, +%u) exceed %llu", s, r0, cnt, (unsigned long long)total_rows);
        for (uint32_t i = 0; i < cnt; ++i) {
            const int bin = h_row_sym[r0 + i] + ((s % 3) == HIC_KIND_LENGTH ? 0 : g.nb_bins / 2);
            HIC_REQUIRE(bin >= 0 && bin < g.nb_bins, "stream %d: symbol %d outside the plan's bins", s, h_row_sym[r0 + i]);
            row_stream[r0 + i] = (uint32_t)s;
        }
        p->rows[s] = cnt;
        p->dev_row_start[s] = r0;
        p->row_off[s + 1] = p->row_off[s] + cnt;
        p->nsym[s] = h_nsym[s];
        p->nbits[s] = h_nbits[s];
        if (h_start_bit) {
            HIC_REQUIRE(h_start_bit[s] <= 8, "stream %d: start bit %u", s, h_start_bit[s]);
            start[s] = h_start_bit[s];
        }
        if (h_nsym[s] == 0) continue;
        const uint64_t len = start[s] == 8 ? 1 + (h_nbits[s] + (8 - (h_nbits[s] & 7))) / 8 : (start[s] + h_nbits[s] + 7) / 8;
        p->byte_off[s] = off;
        p->byte_len[s] = len;
        off += (len + 3) & ~3ull;
    }
    p->total_bytes = off;
    p->total_rows = total_rows;
    int rc = ensure_row_capacity(p, total_rows + 1);
    if (rc) return rc;
    HIC_CUDA(cudaMemcpyAsync(p->d_row_sym, h_row_sym, sizeof(int32_t) * total_rows, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_row_code, h_row_packed, sizeof(uint64_t) * total_rows, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_row_stream, row_stream.data(), sizeof(uint32_t) * total_rows, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_nsym, p->nsym.data(), sizeof(uint32_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_nbits, p->nbits.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_byte_off, p->byte_off.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_byte_len, p->byte_len.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_start_bit, start.data(), sizeof(uint32_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_index, h_index, sizeof(CompactIndex) * nss, cudaMemcpyHostToDevice, st));
    if (total_rows)
        HIC_LAUNCH("lut_scatter_kernel", st, lut_scatter_kernel<<<(unsigned)((total_rows + 255) / 256), 256, 0, st>>>(g, p->d_row_sym, p->d_row_code,
                                                                                   p->d_row_stream, total_rows, p->d_lut, p->d_lut_len));
    HIC_CUDA(cudaStreamSynchronize(st));      // the staging vectors go out of scope
    p->start_is_default = h_start_bit == nullptr;
    p->codes_ready = true;
    p->device_built = false;
    p->host_info_valid = true;
    p->host_tables_valid = false;
    return HIC_OK;
}

int hic_entropy_build_codes(hic_entropy_plan* p, void* stream) {
    HIC_REQUIRE(p != nullptr, "plan is NULL");
    const Geom& g = p->g;
    cudaStream_t st = as_stream(stream);
    p->last_stream = st;
    p->prefer_device = false;
    {
        int rc0 = join_dc_pass(p, st);
        if (rc0) return rc0;
        rc0 = restore_default_start(p, st);
        if (rc0) return rc0;
    }
    uint32_t flags[4];
    HIC_CUDA(cudaMemcpyAsync(flags, p->d_err, sizeof(flags), cudaMemcpyDeviceToHost, st));
    std::vector<CompactIndex> index(p->n_ss);
    HIC_CUDA(cudaMemcpyAsync(index.data(), p->d_index, sizeof(CompactIndex) * p->n_ss, cudaMemcpyDeviceToHost, st));
    HIC_CUDA(cudaStreamSynchronize(st));
    if (flags[0]) return hic::fail(HIC_ERR_INVALID, "a symbol fell outside [-%d, %d): create the plan with more value_bins",
                                   g.nb_bins / 2, g.nb_bins / 2);
    const uint64_t n_entries = flags[1];
    std::vector<CompactEntry> entries(n_entries);
    if (n_entries) {
        HIC_CUDA(cudaMemcpyAsync(entries.data(), p->d_entries, sizeof(CompactEntry) * n_entries, cudaMemcpyDeviceToHost, st));
        HIC_CUDA(cudaStreamSynchronize(st));
    }
    const int nss = p->n_ss;
    p->rows.assign(nss, 0);
    p->nsym.assign(nss, 0);
    p->nbits.assign(nss, 0);
    p->byte_off.assign(nss, 0);
    p->byte_len.assign(nss, 0);
    p->row_off.assign(nss + 1, 0);
    for (int s = 0; s < nss; ++s) {
        p->rows[s] = index[s].count;
        p->row_off[s + 1] = p->row_off[s] + index[s].count;
    }
    p->total_rows = p->row_off[nss];
    p->t_sym.resize(p->total_rows);
    p->t_len.resize(p->total_rows);
    p->t_code.resize(p->total_rows);
    std::vector<uint32_t> row_stream(p->total_rows);
    std::vector<uint64_t> row_packed(p->total_rows);

    // Huffman construction, one stream at a time per worker (the reference does this serially in Python)
    unsigned hw = std::thread::hardware_concurrency();
    const int workers = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
    std::vector<int> status(workers, 0);
    auto work = [&](int wid) {
        HeapqHuffman huff;
        std::vector<CompactEntry> local;
        std::vector<uint32_t> freqs;
        std::vector<HuffCode> codes;
        for (int s = wid; s < nss; s += workers) {
            const CompactIndex ix = index[s];
            if (!ix.count) continue;
            local.assign(entries.begin() + ix.offset, entries.begin() + ix.offset + ix.count);
            std::sort(local.begin(), local.end(), [](const CompactEntry& a, const CompactEntry& b) { return a.first < b.first; });
            freqs.resize(ix.count);
            for (uint32_t i = 0; i < ix.count; ++i) freqs[i] = local[i].count;
            if (!huff.build(freqs.data(), ix.count, codes, MAX_CODE_LEN)) {
                status[wid] = 1;
                return;
            }
            uint64_t bits = 0, n = 0;
            const uint64_t r0 = p->row_off[s];
            for (uint32_t i = 0; i < ix.count; ++i) {
                p->t_sym[r0 + i] = local[i].sym;
                p->t_len[r0 + i] = (uint8_t)codes[i].len;
                p->t_code[r0 + i] = codes[i].bits;
                row_stream[r0 + i] = (uint32_t)s;
                row_packed[r0 + i] = ((uint64_t)codes[i].len << 58) | codes[i].bits;
                bits += (uint64_t)local[i].count * codes[i].len;
                n += local[i].count;
            }
            p->nbits[s] = bits;
            p->nsym[s] = (uint32_t)n;
        }
    };
    if (workers == 1 || nss < 16) {
        for (int wdx = 0; wdx < workers; ++wdx) work(wdx);
    } else {
        std::vector<std::thread> pool;
        for (int wdx = 0; wdx < workers; ++wdx) pool.emplace_back(work, wdx);
        for (auto& t : pool) t.join();
    }
    for (int wdx = 0; wdx < workers; ++wdx)
        if (status[wdx]) return hic::fail(HIC_ERR_INVALID, "a Huffman code exceeds %u bits", MAX_CODE_LEN);

    uint64_t off = 0;
    for (int s = 0; s < nss; ++s) {
        if (p->nsym[s] == 0) continue;                       // no such stream (DC in flat mode)
        const uint64_t pad = 8 - (p->nbits[s] & 7);
        p->byte_off[s] = off;
        p->byte_len[s] = 1 + (p->nbits[s] + pad) / 8;
        off += (p->byte_len[s] + 3) & ~3ull;
    }
    p->total_bytes = off;

    if (p->total_rows > p->row_capacity) {
        if (p->d_row_sym) cudaFree(p->d_row_sym);
        if (p->d_row_code) cudaFree(p->d_row_code);
        if (p->d_row_stream) cudaFree(p->d_row_stream);
        p->d_row_sym = nullptr; p->d_row_code = nullptr; p->d_row_stream = nullptr;
        p->row_capacity = p->total_rows + p->total_rows / 4 + 1024;
        HIC_CUDA(dalloc(&p->d_row_sym, p->row_capacity));
        HIC_CUDA(dalloc(&p->d_row_code, p->row_capacity));
        HIC_CUDA(dalloc(&p->d_row_stream, p->row_capacity));
    }
    HIC_CUDA(cudaMemcpyAsync(p->d_row_sym, p->t_sym.data(), sizeof(int32_t) * p->total_rows, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_row_code, row_packed.data(), sizeof(uint64_t) * p->total_rows, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_row_stream, row_stream.data(), sizeof(uint32_t) * p->total_rows, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_nsym, p->nsym.data(), sizeof(uint32_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_nbits, p->nbits.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_byte_off, p->byte_off.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    if (p->total_rows) {
        HIC_LAUNCH("lut_scatter_kernel", st, lut_scatter_kernel<<<(unsigned)((p->total_rows + 255) / 256), 256, 0, st>>>(g, p->d_row_sym, p->d_row_code,
                                                                                   p->d_row_stream, p->total_rows, p->d_lut, p->d_lut_len));
    }
    HIC_CUDA(cudaMemcpyAsync(p->d_ss_byte_len, p->byte_len.data(), sizeof(uint64_t) * nss, cudaMemcpyHostToDevice, st));
    // rows were uploaded in stream order: make the device index describe that layout
    std::vector<CompactIndex> row_index(nss);
    for (int s = 0; s < nss; ++s) row_index[s] = CompactIndex{(uint32_t)p->row_off[s], p->rows[s]};
    HIC_CUDA(cudaMemcpyAsync(p->d_index, row_index.data(), sizeof(CompactIndex) * nss, cudaMemcpyHostToDevice, st));
    HIC_CUDA(cudaStreamSynchronize(st));      // the staging vectors above go out of scope
    p->codes_ready = true;
    p->device_built = false;
    p->host_info_valid = true;
    p->host_tables_valid = true;
    return HIC_OK;
}

static int ensure_row_capacity(hic_entropy_plan* p, uint64_t rows) {
    if (rows <= p->row_capacity) return HIC_OK;
    if (p->d_row_sym) cudaFree(p->d_row_sym);
    if (p->d_row_code) cudaFree(p->d_row_code);
    if (p->d_row_stream) cudaFree(p->d_row_stream);
    p->d_row_sym = nullptr; p->d_row_code = nullptr; p->d_row_stream = nullptr;
    p->row_capacity = rows;
    HIC_CUDA(dalloc(&p->d_row_sym, p->row_capacity));
    HIC_CUDA(dalloc(&p->d_row_code, p->row_capacity));
    HIC_CUDA(dalloc(&p->d_row_stream, p->row_capacity));
    return HIC_OK;
}

// mirror the per-stream results of a device build on the host (small: 32 bytes per stream)
static int fetch_host_info(hic_entropy_plan* p, cudaStream_t st) {
    if (p->host_info_valid) return HIC_OK;
    const int nss = p->n_ss;
    std::vector<CompactIndex> index(nss);
    p->rows.assign(nss, 0); p->nsym.assign(nss, 0); p->nbits.assign(nss, 0);
    p->byte_off.assign(nss, 0); p->byte_len.assign(nss, 0); p->row_off.assign(nss + 1, 0);
    {
        const void* h[5] = {};
        const void* src[5] = {p->d_index, p->d_ss_nsym, p->d_ss_nbits, p->d_ss_byte_off, p->d_ss_byte_len};
        void* dst[5] = {index.data(), p->nsym.data(), p->nbits.data(), p->byte_off.data(), p->byte_len.data()};
        const size_t bytes[5] = {sizeof(CompactIndex) * nss, sizeof(uint32_t) * nss, sizeof(uint64_t) * nss, sizeof(uint64_t) * nss,
                                 sizeof(uint64_t) * nss};
        p->xfer.reset();
        for (int i = 0; i < 5; ++i) {
            int rc = hic::small_d2h(p->xfer, src[i], bytes[i], st, &h[i]);
            if (rc) return rc;
        }
        HIC_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < 5; ++i) memcpy(dst[i], h[i], bytes[i]);
        p->xfer.reset();
    }
    // rows of a device build sit at the compaction offsets, which are not in stream order
    p->dev_row_start.assign(nss, 0);
    for (int s = 0; s < nss; ++s) {
        p->rows[s] = index[s].count;
        p->dev_row_start[s] = index[s].offset;
        p->row_off[s + 1] = p->row_off[s] + index[s].count;
    }
    p->host_info_valid = true;
    return HIC_OK;
}

static int fetch_host_tables(hic_entropy_plan* p, cudaStream_t st) {
    if (p->host_tables_valid) return HIC_OK;
    int rc = fetch_host_info(p, st);
    if (rc) return rc;
    std::vector<int32_t> sym(p->total_rows);
    std::vector<uint64_t> code(p->total_rows);
    if (p->total_rows) {
        HIC_CUDA(cudaMemcpyAsync(sym.data(), p->d_row_sym, sizeof(int32_t) * p->total_rows, cudaMemcpyDeviceToHost, st));
        HIC_CUDA(cudaMemcpyAsync(code.data(), p->d_row_code, sizeof(uint64_t) * p->total_rows, cudaMemcpyDeviceToHost, st));
        HIC_CUDA(cudaStreamSynchronize(st));
    }
    p->t_sym.resize(p->total_rows); p->t_len.resize(p->total_rows); p->t_code.resize(p->total_rows);
    for (int s = 0; s < p->n_ss; ++s) {
        const uint64_t src = p->dev_row_start[s], dst = p->row_off[s];
        for (uint32_t i = 0; i < p->rows[s]; ++i) {
            p->t_sym[dst + i] = sym[src + i];
            p->t_len[dst + i] = (uint8_t)(code[src + i] >> 58);
            p->t_code[dst + i] = code[src + i] & ((1ull << 58) - 1);
        }
    }
    p->host_tables_valid = true;
    return HIC_OK;
}

extern "C" int hic_entropy_build_codes_device(hic_entropy_plan* p, void* stream) {
    HIC_REQUIRE(p != nullptr, "plan is NULL");
    const Geom& g = p->g;
    HIC_REQUIRE(g.nb_bins <= 8192, "the device Huffman builder handles up to 8192 value bins; use hic_entropy_build_codes");
    cudaStream_t st = as_stream(stream);
    p->last_stream = st;
    p->prefer_device = true;
    int rc = restore_default_start(p, st);
    if (rc) return rc;
    rc = builder_prepare(p);
    if (rc) return rc;
    uint32_t* leaf_freq = p->d_leaf_freq;
    uint16_t* parent = reinterpret_cast<uint16_t*>(p->d_parent);         // 2 links per entry = 4 bytes
    constexpr int N_AUX = hic_entropy_plan::N_AUX;
    static_assert(DC_LANES < N_AUX, "the run-length pass needs lanes of its own");
    // The replay is latency-bound, so the tiers run side by side on the plan's auxiliary streams.  The DC
    // streams (the longest chains) do not wait for the run-length symbols at all: their pass hangs on the
    // event the emit pass recorded after the DC compaction and overlaps rle_emit on `st` -- the emit pass
    // has already launched it when the previous batch was built on the device too.
    // HIC_ENTROPY_SERIAL (environment): keep the DC pass on `st`, ahead of the run-length pass, so that
    // per-kernel timings are not blurred by the overlap (bench.py's kernel table); results are identical.
    const bool has_dc = p->dc_early && g.L.skip_first;
    const bool serial = serial_build() && !p->dc_launched;
    if (has_dc && serial) {
        rc = builder_sort_pass(p, SEL_DC, p->d_tier_count_dc, p->d_tier_list_dc, st);
        if (rc) return rc;
        hic::prof_begin("huffman_replay_kernel", st);
        cudaStream_t lanes[N_AUX + 1];
        lanes[0] = st;
        HIC_CUDA(cudaEventRecord(p->ev_fork, st));
        for (int a = 0; a < N_AUX; ++a) {
            HIC_CUDA(cudaStreamWaitEvent(p->aux[a], p->ev_fork, 0));
            lanes[a + 1] = p->aux[a];
        }
        rc = builder_replay_pass(p, p->d_tier_count_dc, p->d_tier_list_dc, lanes, N_AUX + 1);
        if (rc) return rc;
        for (int a = 0; a < N_AUX; ++a) {
            HIC_CUDA(cudaEventRecord(p->ev_join[a], p->aux[a]));
            HIC_CUDA(cudaStreamWaitEvent(st, p->ev_join[a], 0));
        }
        hic::prof_end(st);
    } else if (has_dc && !p->dc_launched) {
        rc = launch_dc_pass(p);
        if (rc) return rc;
    }
    const int n_dc_lanes = p->dc_launched ? DC_LANES : 0;
    rc = builder_sort_pass(p, has_dc ? SEL_AC : SEL_ALL, p->d_tier_count, p->d_tier_list, st);
    if (rc) return rc;
    // (profiled as ONE span on `st` from fork to join: the tier launches overlap each other)
    hic::prof_begin("huffman_replay_kernel", st);
    {
        cudaStream_t lanes[N_AUX + 1];
        int n_lanes = 0;
        lanes[n_lanes++] = st;
        HIC_CUDA(cudaEventRecord(p->ev_fork, st));
        for (int a = n_dc_lanes; a < N_AUX; ++a) {
            HIC_CUDA(cudaStreamWaitEvent(p->aux[a], p->ev_fork, 0));
            lanes[n_lanes++] = p->aux[a];
        }
        rc = builder_replay_pass(p, p->d_tier_count, p->d_tier_list, lanes, n_lanes);
        if (rc) return rc;
    }
    for (int a = 0; a < N_AUX; ++a) {
        HIC_CUDA(cudaEventRecord(p->ev_join[a], p->aux[a]));
        HIC_CUDA(cudaStreamWaitEvent(st, p->ev_join[a], 0));
    }
    p->dc_launched = false;
    hic::prof_end(st);
    HIC_LAUNCH("huffman_codes_kernel", st, huffman_codes_kernel<<<p->n_ss, 128, 0, st>>>(
        g, p->d_index, leaf_freq, parent, p->d_row_sym, p->d_row_code, p->d_lut, p->d_lut_len, p->d_ss_nsym, p->d_ss_nbits, p->d_err));
    HIC_LAUNCH("payload_layout_kernel", st, payload_layout_kernel<<<1, 1024, 0, st>>>(p->n_ss, p->d_ss_nsym, p->d_ss_nbits,
        p->d_ss_byte_off, p->d_ss_byte_len, p->d_pay_totals));
    unsigned long long totals[2] = {0, 0};
    uint32_t flags[4];
    {
        // (through the SMs into page-locked staging: a copy-engine transfer would wait behind the bulk
        // downloads other streams have queued)
        const void *h_tot = nullptr, *h_flg = nullptr;
        p->xfer.reset();
        rc = hic::small_d2h(p->xfer, p->d_pay_totals, sizeof(unsigned long long), st, &h_tot);
        if (rc) return rc;
        rc = hic::small_d2h(p->xfer, p->d_err, sizeof(flags), st, &h_flg);
        if (rc) return rc;
        HIC_CUDA(cudaStreamSynchronize(st));
        memcpy(totals, h_tot, sizeof(unsigned long long));
        memcpy(flags, h_flg, sizeof(flags));
        p->xfer.reset();
    }
    if (flags[0] & 1u) return hic::fail(HIC_ERR_INVALID, "a symbol fell outside [-%d, %d): create the plan with more value_bins",
                                        g.nb_bins / 2, g.nb_bins / 2);
    if (flags[0] & 2u) return hic::fail(HIC_ERR_INVALID, "an alphabet exceeds 8192 symbols; use hic_entropy_build_codes");
    if (flags[0] & 4u) return hic::fail(HIC_ERR_INVALID, "a Huffman code exceeds %u bits", MAX_CODE_LEN);
    p->total_bytes = totals[0];
    p->total_rows = flags[1];
    p->codes_ready = true;
    p->device_built = true;
    p->host_info_valid = false;
    p->host_tables_valid = false;
    return HIC_OK;
}

int hic_entropy_stream_info(hic_entropy_plan* p, uint32_t* h_rows, uint32_t* h_nsym, uint64_t* h_nbits,
                            uint64_t* h_byte_off, uint64_t* h_byte_len, uint64_t* total_rows, uint64_t* total_bytes) {
    HIC_REQUIRE(p != nullptr, "plan is NULL");
    HIC_REQUIRE(p->codes_ready, "hic_entropy_build_codes has not run");
    {
        int rc = fetch_host_info(p, p->last_stream);
        if (rc) return rc;
    }
    for (int s = 0; s < p->n_ss; ++s) {
        if (h_rows) h_rows[s] = p->rows[s];
        if (h_nsym) h_nsym[s] = p->nsym[s];
        if (h_nbits) h_nbits[s] = p->nbits[s];
        if (h_byte_off) h_byte_off[s] = p->byte_off[s];
        if (h_byte_len) h_byte_len[s] = p->byte_len[s];
    }
    if (total_rows) *total_rows = p->total_rows;
    if (total_bytes) *total_bytes = p->total_bytes;
    return HIC_OK;
}

int hic_entropy_tables(hic_entropy_plan* p, int32_t* h_symbols, uint8_t* h_lens, uint64_t* h_codes) {
    HIC_REQUIRE(p != nullptr, "plan is NULL");
    HIC_REQUIRE(p->codes_ready, "hic_entropy_build_codes has not run");
    {
        int rc = fetch_host_tables(p, p->last_stream);
        if (rc) return rc;
    }
    for (uint64_t i = 0; i < p->total_rows; ++i) {
        if (h_symbols) h_symbols[i] = p->t_sym[i];
        if (h_lens) h_lens[i] = p->t_len[i];
        if (h_codes) h_codes[i] = p->t_code[i];
    }
    return HIC_OK;
}

int hic_entropy_tables_packed(hic_entropy_plan* p, uint32_t* h_index, int32_t* h_row_sym, uint64_t* h_row_packed,
                              void* stream) {
    HIC_REQUIRE(p && h_index && h_row_sym && h_row_packed, "NULL argument");
    HIC_REQUIRE(p->codes_ready, "hic_entropy_build_codes has not run");
    cudaStream_t st = as_stream(stream);
    static_assert(sizeof(CompactIndex) == 2 * sizeof(uint32_t), "index layout");
    int rc = hic::small_d2h_to(h_index, p->d_index, sizeof(CompactIndex) * p->n_ss, st);
    if (rc) return rc;
    if (p->total_rows) {
        rc = hic::small_d2h_to(h_row_sym, p->d_row_sym, sizeof(int32_t) * p->total_rows, st);
        if (rc) return rc;
        rc = hic::small_d2h_to(h_row_packed, p->d_row_code, sizeof(uint64_t) * p->total_rows, st);
        if (rc) return rc;
    }
    return HIC_OK;
}

int hic_entropy_pack(hic_entropy_plan* p, uint8_t* d_out, void* stream) {
    HIC_REQUIRE(p && d_out, "NULL argument");
    HIC_REQUIRE(p->codes_ready, "hic_entropy_build_codes has not run");
    HIC_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 3) == 0, "d_out must be 4-byte aligned");
    const Geom& g = p->g;
    cudaStream_t st = as_stream(stream);
    HIC_CUDA(cudaMemsetAsync(d_out, 0, p->total_bytes, st));
    const unsigned tiles = (unsigned)p->total_ptiles;
    HIC_LAUNCH("pack_tile_bits_kernel", st, pack_tile_bits_kernel<<<tiles, PACK_THREADS, 0, st>>>(g, p->d_lut_len, p->d_ss_nsym, p->d_dc, p->d_values, p->d_lengths,
                                                         p->d_ptile_bits));
    HIC_LAUNCH("pack_stream_scan_kernel", st, pack_stream_scan_kernel<<<(p->n_ss + 127) / 128, 128, 0, st>>>(g, p->d_ptile_bits, p->d_start_bit, p->d_ptile_off));
    HIC_LAUNCH("pack_emit_kernel", st, pack_emit_kernel<<<tiles, PACK_THREADS, 0, st>>>(g, p->d_lut, p->d_ss_nsym, p->d_ss_nbits, p->d_ss_byte_off,
                                                    p->d_ptile_off, p->d_start_bit, p->d_dc, p->d_values, p->d_lengths, d_out));
    return HIC_OK;
}

int hic_entropy_device_tables(const hic_entropy_plan* p, const void** d_index, const int32_t** d_row_sym,
                              const uint64_t** d_row_packed, const uint64_t** d_byte_off, const uint64_t** d_nbits) {
    HIC_REQUIRE(p != nullptr, "plan is NULL");
    HIC_REQUIRE(p->codes_ready, "no codes have been built");
    if (d_index) *d_index = p->d_index;
    if (d_row_sym) *d_row_sym = p->d_row_sym;
    if (d_row_packed) *d_row_packed = p->d_row_code;
    if (d_byte_off) *d_byte_off = p->d_ss_byte_off;
    if (d_nbits) *d_nbits = p->d_ss_nbits;
    return HIC_OK;
}

int hic_huffman_build_host(const uint32_t* h_freqs, uint32_t n, uint8_t* h_lens, uint64_t* h_codes) {
    HIC_REQUIRE(h_freqs && h_lens && h_codes, "NULL argument");
    HeapqHuffman huff;
    std::vector<HuffCode> codes;
    if (!huff.build(h_freqs, n, codes, MAX_CODE_LEN)) return hic::fail(HIC_ERR_INVALID, "a Huffman code exceeds %u bits", MAX_CODE_LEN);
    for (uint32_t i = 0; i < n; ++i) {
        h_lens[i] = (uint8_t)codes[i].len;
        h_codes[i] = codes[i].bits;
    }
    return HIC_OK;
}

int hic_entropy_symbol_buffers(const hic_entropy_plan* p, const int16_t** d_dc, const int16_t** d_values,
                               const uint8_t** d_lengths) {
    HIC_REQUIRE(p != nullptr, "plan is NULL");
    if (d_dc) *d_dc = p->d_dc;
    if (d_values) *d_values = p->d_values;
    if (d_lengths) *d_lengths = p->d_lengths;
    return HIC_OK;
}

}  // extern "C"
