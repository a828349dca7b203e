"""A rigorous float32 rounding-error bound for K1's / K7's two-pass 8x8 transforms (eo_forward8 / eo_inverse8
in hiccup_b200/csrc/hic_core.cuh), derived independently of the C++ derivation in csrc/hic_dct_bound.h (the
library's near-tie bands come from that one; tests/test_oracle_dct.py checks that the two agree).

Model.  Every value n of the flow graph is a linear functional of the 64 inputs, n(x) = sum_i c_ni x_i.  A
float32 operation returns the exact result times (1 + d), |d| <= u = 2^-24 (round to nearest; an FMA rounds
once, so a fused multiply-add is never worse than the separate pair bounded here), and a float32 constant is
its exact value times (1 + d).  To first order the error of an output o is  sum_n G_no d_n n(x)  (G_no = the
gain from node n to o), so

    |err_o| <= u sum_n |G_no| |n(x)| <= u sum_i |x_i| (sum_n |G_no| |c_ni|) <= u E max_i sum_n |G_no| |c_ni|,

E = sum_i |x_i| -- the sum of absolute values of linear forms is convex, so its maximum over the L1 ball sits
on a vertex.  The per-input sums are carried forward through the graph as a vector e (with every path gain
taken in absolute value, which only loosens the bound):

    add / sub:   e(a +- b) = e(a) + e(b) + |c(a +- b)|
    mul by k:    e(k a)    = |k| e(a) + 2 |c(k a)|           (product rounding + the constant's own rounding)

and the bound for an output is max_i e_i, in units of u E.  It holds for EVERY input block: no cancellation
between rounding errors is assumed.  Second-order terms are ~(graph depth * u) = 1e-6 of the first-order ones.

Forward (K1): the kernel computes b[u][v] = S_uv * scale_u * scale_v, then t = fma(b, rq, MAGIC) with the float
constant rq = 4 / (scale_u scale_v q).  The reference value is C/q, C = 4 S_uv.  The kernel trusts t when the
distance of b*rq to a half-integer exceeds kappa(u, v) * margin * 4 u E / q, with

    kappa(u, v) = e(b[u][v]) / (|scale_u scale_v| u E)  +  1          (+1: rq's own rounding, |C/q| <= 4 E / q)

Inverse (K7): v = coef * dq (dq = q pre_u pre_v / 256 a float constant; the product is rounded), two inverse
passes, + 128.  The kernel trusts a sample whose distance to an integer exceeds
(4 u / 256) * margin * sum_k w_k |coef_k q_k| + 2^-15 (the last term: the rounding of the final + 128, at most
half an ulp of a value below 512), w_k = the maximum over the 64 samples of input k's error sum.

Why not Arai-Agui-Nakajima (round 1's choice)?  `python tools/dct_error_bound.py --aan` prints the same bound
for it: 224 (forward, coefficient (7,7)) and 337 (inverse) against 12.6 and 9.8 here -- its odd part subtracts
large intermediates.  Round 1's empirical band (16 units) was therefore not a guarantee.

    python tools/dct_error_bound.py [--aan]
"""
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class V:
    """A value of the flow graph: exact linear form c over the inputs + the per-input first-order error sums e
    (units of u; the output bound is max_i e_i per unit E)."""
    __slots__ = ("c", "e")

    def __init__(self, c, e=None):
        self.c = c
        self.e = np.zeros_like(c) if e is None else e

    @property
    def bound(self):
        return float(np.max(self.e))

    def __add__(self, o):
        c = self.c + o.c
        return V(c, self.e + o.e + np.abs(c))

    def __sub__(self, o):
        c = self.c - o.c
        return V(c, self.e + o.e + np.abs(c))

    def scale(self, k):
        c = self.c * k
        return V(c, abs(k) * self.e + 2.0 * np.abs(c))


def fma(k, a, b):
    """k * a + b as ONE fused operation or as mul then add -- bound the worse of the two (the compiler may
    or may not contract): the separate pair."""
    return a.scale(k) + b


T8 = math.tan(math.pi / 8)


def _c(m):
    return math.cos(m * math.pi / 16)


def eo_forward8(x):
    """hic_core.cuh eo_forward8: even / odd split, dense odd part (every row divided by its first entry)."""
    s = [x[i] + x[7 - i] for i in range(4)]
    d = [x[i] - x[7 - i] for i in range(4)]
    a, p, b, q = s[0] + s[3], s[0] - s[3], s[1] + s[2], s[2] - s[1]
    o0, o4 = a + b, a - b
    o2 = fma(-T8, q, p)
    o6 = fma(T8, p, q)
    rows = {1: [_c(1), _c(3), _c(5), _c(7)], 3: [_c(3), -_c(7), -_c(1), -_c(5)],
            5: [_c(5), -_c(1), _c(7), _c(3)], 7: [_c(7), -_c(5), _c(3), -_c(1)]}
    odd = {}
    for k, row in rows.items():
        acc = d[0]
        for i in (1, 2, 3):
            acc = fma(row[i] / row[0], d[i], acc)
        odd[k] = acc
    return [o0, odd[1], o2, odd[3], o4, odd[5], o6, odd[7]]


def eo_inverse8(x):
    """hic_core.cuh eo_inverse8 (inputs prescaled; pivots X1 -> o0, X5 -> o1, X3 -> o2, X7 -> o3)."""
    f13, f15, f17 = _c(3) / _c(1), _c(5) / _c(1), _c(7) / _c(1)
    A, B = x[0] + x[4], x[0] - x[4]
    C = fma(T8, x[6], x[2])
    D = fma(-T8, x[2], x[6])
    e0, e3, e1, e2 = A + C, A - C, B - D, B + D
    o0 = fma(-f17, x[7], fma(-f15, x[5], fma(-f13, x[3], x[1])))
    o1 = fma(f15, x[7], fma(f17, x[3], fma(f13, x[1], x[5])))
    o2 = fma(-f13, x[7], fma(-f17, x[5], fma(f15, x[1], x[3])))
    o3 = fma(-f13, x[5], fma(f15, x[3], fma(f17, x[1], x[7])))
    return [e0 + o0, e1 + o1, e2 + o2, e3 + o3, e3 - o3, e2 - o2, e1 - o1, e0 - o0]


def aan_forward8(d):
    """Round 1's Arai-Agui-Nakajima forward transform (for comparison only)."""
    d0, d1, d2, d3, d4, d5, d6, d7 = d
    t0, t7, t1, t6 = d0 + d7, d0 - d7, d1 + d6, d1 - d6
    t2, t5, t3, t4 = d2 + d5, d2 - d5, d3 + d4, d3 - d4
    e10, e13, e11, e12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    o0, o4 = e10 + e11, e10 - e11
    z1 = (e12 + e13).scale(0.70710678118654752440)
    o2, o6 = e13 + z1, e13 - z1
    o10, o11, o12 = t4 + t5, t5 + t6, t6 + t7
    z5 = (o10 - o12).scale(0.38268343236508977173)
    z2 = fma(0.54119610014619698440, o10, z5)
    z4 = fma(1.30656296487637652786, o12, z5)
    z3 = o11.scale(0.70710678118654752440)
    z11, z13 = t7 + z3, t7 - z3
    return [o0, z11 + z4, o2, z13 - z2, o4, z13 + z2, o6, z11 - z4]


def aan_inverse8(d):
    """Round 1's inverse (for comparison only)."""
    d0, d1, d2, d3, d4, d5, d6, d7 = d
    e10, e11, e13 = d0 + d4, d0 - d4, d2 + d6
    e12 = (d2 - d6).scale(1.41421356237309504880) - e13
    t0, t3, t1, t2 = e10 + e13, e10 - e13, e11 + e12, e11 - e12
    z13, z10, z11, z12 = d5 + d3, d5 - d3, d1 + d7, d1 - d7
    t7 = z11 + z13
    o11 = (z11 - z13).scale(1.41421356237309504880)
    z5 = (z10 + z12).scale(1.84775906502257351225)
    o10 = z5 - z12.scale(1.08239220029239396880)
    o12 = z5 - z10.scale(2.61312592975275305571)
    t6 = o12 - t7
    t5 = o11 - t6
    t4 = o10 - t5
    return [t0 + t7, t1 + t6, t2 + t5, t3 + t4, t3 - t4, t2 - t5, t1 - t6, t0 - t7]


def two_pass(fn, grid):
    rows = [fn(list(grid[r])) for r in range(8)]
    cols = [fn([rows[r][c] for r in range(8)]) for c in range(8)]
    return [[cols[c][r] for c in range(8)] for r in range(8)]          # [r][c]


def forward_scales(fn):
    """out[k] = S_k * scale_k: from the unit impulse at n = 0, whose S_k is cos(k pi / 16)."""
    out = fn([V(np.eye(8)[i].copy()) for i in range(8)])
    return [float(out[k].c[0]) / math.cos(k * math.pi / 16) for k in range(8)]


def inverse_prescales(fn):
    """in[k] = X_k * pre_k so that out[0] has X_k with coefficient 1 (k = 0) or 2 cos(k pi / 16)."""
    out = fn([V(np.eye(8)[i].copy()) for i in range(8)])
    return [(1.0 if k == 0 else 2.0 * math.cos(k * math.pi / 16)) / float(out[0].c[k]) for k in range(8)]


def forward_kappa(fn=None):
    """kappa(u, v) for every coefficient; the inputs x - 128 are exact small integers (no input error)."""
    fn = fn or eo_forward8
    sc = forward_scales(fn)
    eye = np.eye(64)
    grid = [[V(eye[8 * r + c].copy()) for c in range(8)] for r in range(8)]
    out = two_pass(fn, grid)
    kap = np.zeros((8, 8))
    for u in range(8):
        for v in range(8):
            b = out[u][v]
            want = np.array([[math.cos((2 * i + 1) * u * math.pi / 16) * math.cos((2 * j + 1) * v * math.pi / 16)
                              for j in range(8)] for i in range(8)]).reshape(-1) * sc[u] * sc[v]
            assert np.allclose(b.c, want, atol=1e-12), (u, v)
            kap[u, v] = b.bound / abs(sc[u] * sc[v]) + 1.0
    return kap


def inverse_weights(fn=None):
    """w[u, v]: the maximum over the 64 samples of input (u, v)'s error sum, in units of 4 u / 256 per |coef q|;
    also the per-sample maximum over inputs (what a single uniform kappa would have to be)."""
    fn = fn or eo_inverse8
    pre = inverse_prescales(fn)
    eye = np.eye(64)
    grid = [[None] * 8 for _ in range(8)]
    for u in range(8):
        for v in range(8):
            c = eye[8 * u + v] * (pre[u] * pre[v] / 256.0)
            grid[u][v] = V(c, 2.0 * np.abs(c))
    out = two_pass(fn, grid)
    e = np.zeros((64, 64))                                  # [input][sample]
    for y in range(8):
        for x in range(8):
            p = out[y][x]
            want = np.array([[(1.0 if uu == 0 else 2.0 * math.cos((2 * y + 1) * uu * math.pi / 16)) *
                              (1.0 if vv == 0 else 2.0 * math.cos((2 * x + 1) * vv * math.pi / 16)) / 256.0
                              for vv in range(8)] for uu in range(8)]).reshape(-1)
            assert np.allclose(p.c, want, atol=1e-12), (y, x)
            e[:, 8 * y + x] = p.e * 256.0 / 4.0
    return e.max(axis=1).reshape(8, 8), e.max(axis=0).reshape(8, 8)


def main():
    import sys
    aan = "--aan" in sys.argv
    fk = forward_kappa(aan_forward8 if aan else eo_forward8)
    w, per_sample = inverse_weights(aan_inverse8 if aan else eo_inverse8)
    fk_ac = fk.copy()
    fk_ac[0, 0] = 0.0                                   # DC is an exact integer (4 * sum x) and is never flagged
    np.set_printoptions(precision=2, suppress=True, linewidth=140)
    print("%s transforms" % ("Arai-Agui-Nakajima (round 1)" if aan else "even/odd with dense odd part (hic_core.cuh)"))
    print("forward kappa(u, v) (rigorous first-order bound, units of 4 u E / q):")
    print(fk)
    print("max over AC coefficients: %.3f" % fk_ac.max())
    print("inverse weights w(u, v) per input coefficient (units of 4 u / 256 per |coef q|):")
    print(w)
    print("max: %.3f   (per-sample maximum over inputs: %.3f)" % (w.max(), per_sample.max()))


if __name__ == "__main__":
    main()
