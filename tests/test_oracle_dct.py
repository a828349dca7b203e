"""The oracle's float64 DCT restatement against scipy (ducc0), and the C restatement against both."""
import ctypes
import os

import numpy as np
import pytest
import scipy.fftpack

from oracle import hiccup_oracle as orc

HARNESS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_harness", "libhic_cpu_harness.so")


def _harness():
    if not os.path.exists(HARNESS):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(HARNESS)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_dct2_bit_identical_to_scipy():
    rng = np.random.default_rng(0)
    for lo, hi in ((-128, 128), (-32768, 32768), (-3, 4)):
        x = rng.integers(lo, hi, (20000, 8)).astype(np.float64)
        assert np.array_equal(orc.ducc_dct2_8(x), scipy.fftpack.dct(x, axis=1))
    x = rng.standard_normal((20000, 8)) * 1000
    assert np.array_equal(orc.ducc_dct2_8(x), scipy.fftpack.dct(x, axis=1))


def test_dct3_bit_identical_to_scipy():
    rng = np.random.default_rng(1)
    x = rng.integers(-200000, 200000, (20000, 8)).astype(np.float64)
    assert np.array_equal(orc.ducc_dct3_8(x), scipy.fftpack.idct(x, axis=1))
    x = rng.standard_normal((20000, 8)) * 1000
    assert np.array_equal(orc.ducc_dct3_8(x), scipy.fftpack.idct(x, axis=1))


def test_row_by_row_calls_equal_axis_calls():
    """The reference transforms one row at a time (transform.py:79-82); the oracle uses whole arrays."""
    rng = np.random.default_rng(2)
    x = rng.integers(-128, 128, (64, 8)).astype(np.float64)
    rows = np.array([scipy.fftpack.dct(r) for r in x])
    assert np.array_equal(rows, scipy.fftpack.dct(x, axis=1))


def test_dct_identity_times_constant():
    # transformtest.py:55-68: idct2(dct2(x)) == x (the /256 is inside idct2)
    rng = np.random.default_rng(3)
    b = rng.integers(-128, 128, (10, 8, 8)).astype(np.float64)
    assert np.allclose(orc.idct2_blocks(orc.dct2_blocks(b)), b, atol=1e-9)


def test_c_restatement_equals_oracle():
    """csrc/hic_core.cuh (ducc_dct2_8 / ducc_dct3_8 with embedded hex twiddles) == the oracle, which
    derives the twiddles at run time the way ducc0 does."""
    lib = _harness()
    rng = np.random.default_rng(4)
    x = rng.integers(-128, 128, (50000, 8)).astype(np.float64)
    a = x.copy()
    lib.hx_ducc_dct2(_ptr(a), len(a))
    assert np.array_equal(a, orc.ducc_dct2_8(x))
    x = rng.integers(-200000, 200000, (50000, 8)).astype(np.float64)
    a = x.copy()
    lib.hx_ducc_dct3(_ptr(a), len(a))
    assert np.array_equal(a, orc.ducc_dct3_8(x))


def _bounds():
    lib = _harness()
    kap, w, sc, pre = np.zeros(64), np.zeros(64), np.zeros(8), np.zeros(8)
    lib.hx_dct_bounds(_ptr(kap), _ptr(w), _ptr(sc), _ptr(pre))
    return kap.reshape(8, 8), w.reshape(8, 8), sc, pre


def test_error_bounds_two_derivations_agree():
    """The near-tie bands are RIGOROUS bounds of the float32 rounding error of eo_forward8 / eo_inverse8.  The
    library derives them from the kernel's own template code run on a bound-carrying value type
    (csrc/hic_dct_bound.h); tools/dct_error_bound.py restates transforms and bound in Python.  Same numbers."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dct_error_bound", os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "dct_error_bound.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    kap, w, sc, pre = _bounds()
    assert np.allclose(kap, tool.forward_kappa(), rtol=1e-9)
    assert np.allclose(w, tool.inverse_weights()[0], rtol=1e-9)
    assert np.allclose(sc, tool.forward_scales(tool.eo_forward8), rtol=1e-12)
    assert np.allclose(pre, tool.inverse_prescales(tool.eo_inverse8), rtol=1e-12)
    kap_ac = kap.copy()
    kap_ac[0, 0] = 0
    assert kap_ac.max() < 12.6 and w.max() < 9.8          # what DESIGN.md quotes
    # round 1's Arai-Agui-Nakajima transforms under the same analysis: why they were replaced
    assert tool.forward_kappa(tool.aan_forward8)[7, 7] > 200 and tool.inverse_weights(tool.aan_inverse8)[0].max() > 300


def _forward_ratio(lib, px, kind):
    px = np.ascontiguousarray(px.reshape(-1, 64).astype(np.int16))
    nb = len(px)
    out, mask, ratio = np.zeros((nb, 64), np.int16), np.zeros(nb, np.uint64), np.zeros(nb)
    lib.hx_forward_blocks(_ptr(px), nb, kind, _ptr(out), _ptr(mask), _ptr(ratio))
    return out, mask, ratio


def test_adversarial_search_stays_inside_the_forward_bound(capsys):
    """Look hard for a block whose float32 error approaches the rigorous bound: random blocks of several
    statistics (incl. +-128 checkerboards and sparse spikes, the shapes that maximise sum |x| per coefficient),
    then hill climbing on the worst ones.  The observed maximum must stay below 1 (= the bound); the kernel's
    band is HIC_BAND_MARGIN times the bound on top of that."""
    lib = _harness()
    rng = np.random.default_rng(2024)
    n = 120000
    pools = [rng.integers(-128, 128, (n, 64)),
             rng.choice([-128, 127], (n, 64)),
             (rng.integers(0, 2, (n, 64)) * 255 - 128) * (rng.random((n, 64)) < 0.1),
             np.clip(np.rint(rng.normal(0, 40, (n, 1)) + rng.normal(0, 6, (n, 64))), -128, 127),
             ((np.indices((8, 8)).sum(axis=0) & 1).reshape(-1)[None, :] * 255 - 128) * rng.choice([-1, 1], (n, 1)) + rng.integers(-1, 2, (n, 64))]
    best = 0.0
    for kind in (0, 1):
        for pool in pools:
            px = np.clip(pool, -128, 127).astype(np.int16)
            _, _, ratio = _forward_ratio(lib, px, kind)
            best = max(best, float(ratio.max()))
            # hill climb from the 64 worst blocks of the pool: perturb a few samples, keep improvements
            cur = px[np.argsort(ratio)[-64:]].copy()
            cur_r = np.sort(ratio)[-64:].copy()
            for _ in range(150):
                cand = cur.copy()
                idx = rng.integers(0, 64, (64, 3))
                cand[np.arange(64)[:, None], idx] = rng.integers(-128, 128, (64, 3))
                _, _, r = _forward_ratio(lib, cand, kind)
                better = r > cur_r
                cur[better], cur_r[better] = cand[better], r[better]
            best = max(best, float(cur_r.max()))
    with capsys.disabled():
        print("\n[forward] worst observed float32 error / rigorous bound over 1.2 M blocks + hill climbing: %.3f" % best)
    assert best < 1.0


@pytest.mark.parametrize("kind", [0, 1])
def test_float32_forward_ties_are_all_flagged(kind):
    """K1's arithmetic on the host: every float32 result that differs from the reference's float64
    result lies inside the flagged band, and the observed error stays under the rigorous bound of its
    coefficient (ratio < 1)."""
    lib = _harness()
    rng = np.random.default_rng(5 + kind)
    table = orc.LUM_TABLE if kind == 0 else orc.CHROMA_TABLE
    planes = [orc.rgb_to_ycrcb(orc.synthetic_image(256, 256, 9))[0], rng.integers(0, 256, (128, 128)).astype(np.uint8),
              (rng.integers(0, 2, (128, 128)) * 255).astype(np.uint8)]
    for plane in planes:
        px = np.ascontiguousarray(orc.split_blocks(plane.astype(np.int64) - 128).reshape(-1, 64).astype(np.int16))
        nb = len(px)
        out = np.zeros((nb, 64), np.int16)
        mask = np.zeros(nb, np.uint64)
        ratio = np.zeros(nb)
        lib.hx_forward_blocks(_ptr(px), nb, kind, _ptr(out), _ptr(mask), _ptr(ratio))
        ref = np.ascontiguousarray(orc.blocks_zigzag(orc.dct_channel(plane, table)))
        flagged = ((mask[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
        assert not ((out != ref) & ~flagged).any()
        assert ratio.max() < 1.0
        # the float64 path used by the fix-up kernel reproduces the reference exactly
        q = table.reshape(-1)
        for b, k in zip(*np.nonzero(flagged)):
            nat = int(orc.ZIGZAG8[k])
            exact = lib.hx_exact_coef(_ptr(px[b]), nat // 8, nat % 8, int(q[nat]))
            assert exact == ref[b, k]


@pytest.mark.parametrize("kind", [0, 1])
def test_float32_inverse_near_integers_are_all_flagged(kind):
    lib = _harness()
    table = orc.LUM_TABLE if kind == 0 else orc.CHROMA_TABLE
    ch = "lum" if kind == 0 else "cr"
    worst = 0.0
    for seed in (11, 12):
        planes = orc.jpeg_compression(orc.synthetic_image(128, 192, seed))
        zz = np.ascontiguousarray(orc.blocks_zigzag(planes[ch]).astype(np.int16))
        nb = len(zz)
        out = np.zeros((nb, 64), np.uint8)
        exact = np.zeros((nb, 64), np.uint8)
        mask = np.zeros(nb, np.uint64)
        ratio = np.zeros(nb)
        lib.hx_inverse_blocks(_ptr(zz), nb, kind, _ptr(out), _ptr(mask), _ptr(ratio), _ptr(exact))
        ref = orc.split_blocks(orc.inv_dct_channel(planes[ch], table)).reshape(nb, 64)
        flagged = ((mask[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
        assert np.array_equal(exact, ref)
        assert not ((out != ref) & ~flagged).any()
        worst = max(worst, float(ratio.max()))
    assert worst < 1.0


def test_adversarial_search_stays_inside_the_inverse_bound(capsys):
    """The same for K7: random and spiky coefficient blocks (single large coefficients maximise one input's
    share of the band), then hill climbing."""
    lib = _harness()
    rng = np.random.default_rng(4048)
    n = 60000

    def run(zz, kind):
        zz = np.ascontiguousarray(zz.reshape(-1, 64).astype(np.int16))
        nb = len(zz)
        out, exact = np.zeros((nb, 64), np.uint8), np.zeros((nb, 64), np.uint8)
        mask, ratio = np.zeros(nb, np.uint64), np.zeros(nb)
        lib.hx_inverse_blocks(_ptr(zz), nb, kind, _ptr(out), _ptr(mask), _ptr(ratio), _ptr(exact))
        flagged = ((mask[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
        assert not ((out != exact) & ~flagged).any()          # every float32 miss sits inside the band
        return ratio

    best = 0.0
    for kind in (0, 1):
        pools = [rng.integers(-30, 31, (n, 64)),
                 rng.integers(-600, 601, (n, 64)) * (rng.random((n, 64)) < 0.08),
                 rng.integers(-3, 4, (n, 64)) + 900 * (np.arange(64)[None, :] == rng.integers(0, 64, (n, 1)))]
        for pool in pools:
            ratio = run(pool, kind)
            best = max(best, float(ratio.max()))
            cur = pool[np.argsort(ratio)[-64:]].copy()
            cur_r = np.sort(ratio)[-64:].copy()
            for _ in range(100):
                cand = cur.copy()
                idx = rng.integers(0, 64, (64, 2))
                cand[np.arange(64)[:, None], idx] += rng.integers(-40, 41, (64, 2))
                r = run(cand, kind)
                better = r > cur_r
                cur[better], cur_r[better] = cand[better], r[better]
            best = max(best, float(cur_r.max()))
    with capsys.disabled():
        print("\n[inverse] worst observed float32 error / rigorous bound over 360 k blocks + hill climbing: %.3f" % best)
    assert best < 1.0
