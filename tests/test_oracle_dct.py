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


@pytest.mark.parametrize("kind", [0, 1])
def test_float32_forward_ties_are_all_flagged(kind):
    """K1's arithmetic on the host: every float32 result that differs from the reference's float64
    result lies inside the flagged band, and the observed error stays under HIC_TIE_KAPPA."""
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
        assert ratio.max() < 16.0
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
