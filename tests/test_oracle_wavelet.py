"""What the reference's tests pin for the wavelet mode (shapes and perfect reconstruction), checked
on the stand-in transform.  Coefficient VALUES are parity-unpinned (oracle/pywt_standin.py)."""
import numpy as np

from oracle import hiccup_oracle as orc
from oracle import pywt_standin as pywt


def test_level_shapes_50_25_13():
    # transformtest.py:163-168
    dec = pywt.wavedec2(np.zeros((50, 50)), "db1", level=3)
    assert dec[0].shape == (7, 7)
    assert [d[0].shape for d in dec[1:]] == [(7, 7), (13, 13), (25, 25)]


def test_perfect_reconstruction():
    # transformtest.py:170-175
    rng = np.random.default_rng(0)
    for shape in ((64, 64), (48, 80), (50, 50)):
        x = rng.integers(0, 256, shape).astype(np.float64)
        rec = pywt.waverec2(pywt.wavedec2(x, "db1", level=3), "db1")
        assert np.allclose(rec[:shape[0], :shape[1]], x, atol=1e-9)


def test_threshold_kat():
    # transformtest.py:177-193: |v| < t -> 0
    rng = np.random.default_rng(1)
    bands = orc.wavelet_channel(rng.integers(0, 256, (32, 32)).astype(np.uint8))
    for b in bands:
        assert ((np.abs(b) >= orc.WAVELET_THRESHOLD) | (b == 0)).all()


def test_round_trip_is_close():
    rgb = orc.synthetic_image(64, 64, 3)
    planes = orc.wavelet_compression(rgb)
    out = orc.wavelet_decompression(orc.wavelet_decode(orc.wavelet_encode(planes)))
    assert out.shape == rgb.shape
    assert np.abs(out.astype(int) - rgb.astype(int)).mean() < 12
