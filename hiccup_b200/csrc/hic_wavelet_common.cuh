// hic_wavelet_common.cuh -- the whole-matrix zigzag of reference transform._zigzag_indices
// (transform.py:106-124) as closed-form positions, shared by the wavelet kernels.
#pragma once
#include <stdint.h>

namespace hic {

// number of zigzag positions before anti-diagonal d of an h x w matrix
__device__ __forceinline__ int64_t diag_start(int d, int h, int w) {
    const int m = min(h, w), M = max(h, w);
    if (d <= m) return (int64_t)d * (d + 1) / 2;
    if (d <= M - 1) return (int64_t)m * (m + 1) / 2 + (int64_t)(d - m) * m;
    const int64_t r = (int64_t)h + w - 1 - d;
    return (int64_t)h * w - r * (r + 1) / 2;
}
// zigzag position of (y, x): even diagonals run with y ascending, odd ones with y descending
__device__ __forceinline__ int64_t zigzag_pos(int y, int x, int h, int w) {
    const int d = x + y;
    const int y_lo = max(0, d - (w - 1)), y_hi = min(d, h - 1);
    return diag_start(d, h, w) + ((d & 1) ? (y_hi - y) : (y - y_lo));
}

}  // namespace hic
