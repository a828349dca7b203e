// hic_replay.cuh -- the heapq replay of the device Huffman builder on packed 32-bit heap entries.
//
// reference hiccup/huffman.py:60-79 builds its trees with heapq over nodes that compare by frequency
// only (huffman.py:249-250); which of two equal-frequency nodes pops first is decided by heapq's sift
// order, so CPython's heapify / heappop / heappush (_siftup / _siftdown) are replayed operation for
// operation (csrc/hic_huffman.cuh is the host restatement the CPU tests pin this one against).
//
// The replay is one dependent chain per symbol stream, one lane each (hic_entropy.cu), so what counts is
// the latency of one heap level and how many heaps fit in shared memory.  When a stream's symbol count is
// below 2^18 a heap entry packs into ONE word, key = freq << 14 | node (node < 2 n <= 16384), which
//   * halves the shared memory per stream (twice as many streams resident),
//   * makes the two children of 1-based position q one aligned 8-byte pair (slots 2q, 2q+1) and its four
//     grandchildren one aligned 16-byte quad (slots 4q .. 4q+3), and moves an entry with one register.
// A heap level costs one dependent chain: shared-memory round trip (~30 cycles) + the compare / select /
// address arithmetic (~9 dependent ALU operations, ~40 cycles) -- measured 69 cycles per level on B200
// with the round-1 kernel, and the longest chain (a luminance DC alphabet of ~1700 leaves: 37 000 levels)
// IS the duration of the pass, because a whole batch's streams run side by side.  Two modes:
//   MODE 1  one level per step, the grandchild quad requested one level ahead (1.31 ms per pass on C2);
//   MODE 2  TWO levels per step while both are complete (4q + 3 <= size): one LDS.64 + one LDS.128 bring
//           children and grandchildren, the three comparisons run side by side, the second decision is a
//           predicate select, the new position is 4q + 2 r0 + r1 -- one round trip and ~6 dependent
//           operations per two levels; the last one or two levels take single steps with bound checks.
// The item that sifts back up is compared first with the value just moved up (a register), so the common
// case -- it stays at the bottom -- needs no load.
// "a.freq < b.freq" on keys is  a < (b & ~0x3FFF).
//
// __host__ __device__ so that tests/cpu_harness can run the same code against HeapqHuffman.
#pragma once
#include <stdint.h>
#include "hic_core.cuh"

namespace hic {

constexpr int REPLAY_ID_BITS = 14;
constexpr uint32_t REPLAY_ID_MASK = (1u << REPLAY_ID_BITS) - 1u;
constexpr uint32_t REPLAY_NARROW_TOTAL = 1u << (32 - REPLAY_ID_BITS);      // symbol counts below this pack into a word
constexpr int REPLAY_NARROW_MAX_LEAVES = 1 << (REPLAY_ID_BITS - 1);        // 8192

#ifdef __CUDACC__
typedef uint2 rp_pair;
typedef uint4 rp_quad;
#else
struct alignas(8) rp_pair { uint32_t x, y; };
struct alignas(16) rp_quad { uint32_t x, y, z, w; };
#endif

HIC_HD bool rp_less(uint32_t a, uint32_t b) { return a < (b & ~REPLAY_ID_MASK); }

// slot: the stream's heap, 1-based (slot[q], q = 1 .. size; slot[0] unused), 16-byte aligned, with
// `slots` words of room (a multiple of 4, at least n + 4: prefetches past the heap's end read stale words
// that are never used, clamped to the stream's own region).  On entry slot[i + 1] = freq_i << 14 | i for
// the n >= 2 leaves in first-occurrence order.  On exit par[node] = parent | 0x8000 if node is its
// parent's left child (the first popped, bit '1'), for every node but the root 2 n - 2.
template <int MODE>
HIC_HD void replay_narrow(uint32_t* slot, int n, int slots, uint16_t* par) {
    const rp_pair* pair = reinterpret_cast<const rp_pair*>(slot);
    const rp_quad* quad = reinterpret_cast<const rp_quad*>(slot);
    const int quad_lim = slots / 4 - 1;
    int size = n;

    // heapq._siftdown(heap, startpos, pos) with the item in a register
    auto bubble_up = [&](int q0, int q, uint32_t item) {
        while (q > q0) {
            const uint32_t p = slot[q >> 1];
            if (!rp_less(item, p)) break;
            slot[q] = p;
            q >>= 1;
        }
        slot[q] = item;
    };
    // heapq._siftup(heap, pos): the smaller child moves up until a leaf is reached (ties go to the right
    // child), then the item bubbles back up
    auto sift = [&](int q0, uint32_t item) {
        int q = q0;
        uint32_t moved = 0;                       // the value last moved up = slot[q >> 1] once q > q0
        if (MODE == 2) {
            while (4 * q + 3 <= size) {           // both levels complete: no bound checks
                const rp_pair c = pair[q];
                const rp_quad g = quad[q];
                const bool r0 = !rp_less(c.x, c.y);
                const bool rl = !rp_less(g.x, g.y), rr = !rp_less(g.z, g.w);
                const bool r1 = r0 ? rr : rl;
                const uint32_t s0 = r0 ? c.y : c.x;
                const uint32_t ga = r0 ? g.z : g.x, gb = r0 ? g.w : g.y;
                moved = r1 ? gb : ga;
                slot[q] = s0;
                slot[2 * q + (r0 ? 1 : 0)] = moved;
                q = 4 * q + (r0 ? 2 : 0) + (r1 ? 1 : 0);
            }
            while (2 * q <= size) {
                const rp_pair c = pair[q];        // slot 2q+1 may lie past the heap's end (stale, guarded)
                const bool right = (2 * q + 1 <= size) && !rp_less(c.x, c.y);
                moved = right ? c.y : c.x;
                slot[q] = moved;
                q = 2 * q + (right ? 1 : 0);
            }
        } else if (2 * q <= size) {
            rp_pair c = pair[q];
            rp_quad g = quad[q < quad_lim ? q : quad_lim];
            while (true) {
                const bool right = (2 * q + 1 <= size) && !rp_less(c.x, c.y);
                moved = right ? c.y : c.x;
                slot[q] = moved;
                q = 2 * q + (right ? 1 : 0);
                if (2 * q > size) break;
                c.x = right ? g.z : g.x;
                c.y = right ? g.w : g.y;
                g = quad[q < quad_lim ? q : quad_lim];
            }
        }
        if (q > q0 && rp_less(item, moved)) {     // rare: the item climbs
            slot[q] = moved;
            bubble_up(q0, q >> 1, item);
        } else {
            slot[q] = item;
        }
    };
    auto pop = [&]() {
        const uint32_t last = slot[size];
        --size;
        if (size > 0) {
            const uint32_t ret = slot[1];
            sift(1, last);
            return ret;
        }
        return last;
    };

    for (int q = n / 2; q >= 1; --q) sift(q, slot[q]);           // heapq.heapify
    uint32_t next = (uint32_t)n;
    while (size > 1) {
        const uint32_t l = pop();
        const uint32_t r = pop();
        par[l & REPLAY_ID_MASK] = (uint16_t)(next | 0x8000u);
        par[r & REPLAY_ID_MASK] = (uint16_t)next;
        ++size;
        const uint32_t item = (((l >> REPLAY_ID_BITS) + (r >> REPLAY_ID_BITS)) << REPLAY_ID_BITS) | next;
        bubble_up(1, size, item);                                 // heapq.heappush
        ++next;
    }
}

}  // namespace hic
