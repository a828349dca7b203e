"""The oracle's integer colour / pyramid formulas against OpenCV (the reference's third-party
dependency for these steps; opencv 4.13.0 in this image), and the kernels' shared C arithmetic."""
import ctypes
import os

import cv2
import numpy as np
import pytest

from oracle import hiccup_oracle as orc

HARNESS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_harness", "libhic_cpu_harness.so")


def test_rgb_to_ycrcb_matches_cv2():
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    # add the extremes
    rgb[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255], [255, 0, 255]]
    want = cv2.cvtColor(rgb, cv2.COLOR_RGB2YCrCb)
    y, cr, cb = orc.rgb_to_ycrcb(rgb)
    assert np.array_equal(np.stack([y, cr, cb], -1), want)


def test_ycrcb_to_rgb_matches_cv2():
    rng = np.random.default_rng(1)
    ycc = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
    want = cv2.cvtColor(ycc, cv2.COLOR_YCrCb2RGB)
    assert np.array_equal(orc.ycrcb_to_rgb(ycc[..., 0], ycc[..., 1], ycc[..., 2]), want)


@pytest.mark.parametrize("shape", [(64, 64), (33, 47), (26, 40), (2, 2), (5, 9), (426, 640), (7, 2)])
def test_pyr_down_matches_cv2(shape):
    rng = np.random.default_rng(2)
    p = rng.integers(0, 256, shape, dtype=np.uint8)
    want = cv2.pyrDown(p, dstsize=(shape[1] // 2, shape[0] // 2))
    assert np.array_equal(orc.pyr_down(p), want)


@pytest.mark.parametrize("shape", [(32, 32), (16, 23), (13, 20), (1, 1), (2, 5), (213, 320)])
def test_pyr_up_matches_cv2(shape):
    rng = np.random.default_rng(3)
    p = rng.integers(0, 256, shape, dtype=np.uint8)
    want = cv2.pyrUp(p, dstsize=(shape[1] * 2, shape[0] * 2))
    assert np.array_equal(orc.pyr_up(p), want)


def test_pyramids_on_constants():
    # transformtest.py:122-146
    c = np.full((16, 16), 77, np.uint8)
    assert (orc.pyr_down(c) == 77).all() and (orc.pyr_up(c) == 77).all()


def test_kernel_colour_arithmetic_matches_cv2_exhaustively_sampled():
    """hic_core.cuh rgb_to_ycrcb / ycrcb_to_rgb on the host: 2^21 random triples plus all grey/primary ramps."""
    if not os.path.exists(HARNESS):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(HARNESS)
    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 256, (1 << 21, 3), dtype=np.uint8)
    ramp = np.arange(256, dtype=np.uint8)
    extra = np.concatenate([np.stack([ramp, ramp, ramp], 1), np.stack([ramp, 0 * ramp, 0 * ramp], 1),
                            np.stack([0 * ramp, ramp, 0 * ramp], 1), np.stack([0 * ramp, 0 * ramp, ramp], 1),
                            np.stack([ramp, 255 - ramp, ramp // 2], 1)])
    rgb = np.ascontiguousarray(np.concatenate([rgb, extra]))
    out = np.zeros_like(rgb)
    lib.hx_colour(rgb.ctypes.data_as(ctypes.c_void_p), len(rgb), out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(out, cv2.cvtColor(rgb[None], cv2.COLOR_RGB2YCrCb)[0])
    back = np.zeros_like(rgb)
    lib.hx_colour_inv(rgb.ctypes.data_as(ctypes.c_void_p), len(rgb), back.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(back, cv2.cvtColor(rgb[None], cv2.COLOR_YCrCb2RGB)[0])
