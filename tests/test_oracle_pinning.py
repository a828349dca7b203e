"""Pin the oracle: (1) against the committed golden vectors, which were produced by the UNMODIFIED
reference (tests/golden/gen_golden.py); (2) where /root/reference is present (build container
only), against the reference itself run live on fresh inputs."""
import contextlib
import glob
import io
import os
import pickle

import numpy as np
import pytest

from oracle import hiccup_oracle as orc
from tests.conftest import GOLDEN, load_golden

ALL = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))
JPEG = [n for n in ALL if not n.startswith(("w_", "ws_"))]
WAVELET = [n for n in ALL if n.startswith("w_")]


def _stream_content(stream, n_tab):
    import hiccup_b200  # noqa: F401  (resolves the hiccup.hicimage.TupP name inside the pickles)
    from hiccup_b200 import hicimage
    hi = hicimage.HicImage.from_bytes(stream)
    tables = [[(int(a), b) for a, b in p.rows] for p in hi.payloads[:n_tab]]
    bits = [p.byte_stream for p in hi.payloads[n_tab:2 * n_tab]]
    shapes = [tuple(p.numbers) for p in hi.payloads[2 * n_tab:]]
    return tables, bits, shapes


@pytest.mark.parametrize("name", JPEG)
def test_oracle_reproduces_reference_dct_goldens(name):
    g = load_golden(name)
    planes = orc.jpeg_compression(g["rgb"])
    for ch in orc.CHANNELS:
        assert np.array_equal(planes[ch], g["coef_" + ch])
    st = orc.jpeg_streams(planes)
    for ch in orc.CHANNELS:
        assert np.array_equal(st["dc"][ch], g["dc_" + ch])
        assert np.array_equal(st["ac_value"][ch], g["rle_val_" + ch])
        assert np.array_equal(st["ac_length"][ch], g["rle_len_" + ch])
    enc = orc.jpeg_encode(planes)
    tables, bits, shapes = _stream_content(pickle.loads(g["hic"].tobytes()), 9)
    assert [[(int(a), b) for a, b in t] for t in enc["tables"]] == tables
    assert [orc.padded_bits_to_bytes(b) for b in enc["bits"]] == bits
    assert enc["shapes"] == shapes
    if not str(g["decode_error"]):
        dec = orc.jpeg_decode(enc)
        for ch in orc.CHANNELS:
            assert np.array_equal(dec[ch], g["coef_" + ch])
        assert np.array_equal(orc.jpeg_decompression(dec), g["rgb_out"])
    else:
        with pytest.raises(AssertionError):
            orc.jpeg_decode(enc, strict_reference=True)


@pytest.mark.parametrize("name", WAVELET)
def test_oracle_reproduces_reference_wavelet_goldens(name):
    """Entropy stage pinned by the reference's own codec.wavelet_encode; the transform it was fed
    came from the pywt stand-in (PARITY UNPINNED, oracle/pywt_standin.py)."""
    g = load_golden(name)
    planes = orc.wavelet_compression(g["rgb"])
    for ch in orc.CHANNELS:
        for i in range(10):
            assert np.array_equal(planes[ch][i], g["band_%s_%d" % (ch, i)])
    st = orc.wavelet_streams(planes)
    for ch in orc.CHANNELS:
        assert np.array_equal(st["value"][ch], g["rle_val_" + ch])
        assert np.array_equal(st["length"][ch], g["rle_len_" + ch])
    enc = orc.wavelet_encode(planes)
    tables, bits, shapes = _stream_content(pickle.loads(g["hic"].tobytes()), 6)
    assert [[(int(a), b) for a, b in t] for t in enc["tables"]] == tables
    assert [orc.padded_bits_to_bytes(b) for b in enc["bits"]] == bits
    assert enc["shapes"] == shapes
    dec = orc.wavelet_decode(enc)
    for ch in orc.CHANNELS:
        for i in range(10):
            assert np.array_equal(dec[ch][i], planes[ch][i])
    assert np.array_equal(orc.wavelet_decompression(dec), g["rgb_out"])


@pytest.mark.reference
@pytest.mark.parametrize("shape,seed", [((48, 64), 101), ((40, 56), 102), ((80, 80), 103)])
def test_oracle_against_live_reference(shape, seed):
    from oracle import refshim
    refshim.install()
    import hiccup.codec as rcodec
    import hiccup.compression as rcomp
    rgb = orc.synthetic_image(shape[0], shape[1], seed)
    comp = rcomp.jpeg_compression(rgb)
    planes = orc.jpeg_compression(rgb)
    for ch in orc.CHANNELS:
        assert np.array_equal(planes[ch], comp.as_dict[ch])
    hi = rcodec.jpeg_encode(comp)
    enc = orc.jpeg_encode(planes)
    stream = hi.byte_stream()
    for i in range(9):
        assert stream[10 + i] == orc.padded_bits_to_bytes(enc["bits"][i])
        assert [(int(a), b) for a, b in (p.numbers for p in hi.payloads[i].payloads)] == [(int(a), b) for a, b in enc["tables"][i]]
    if shape[0] % 16 == 0 and shape[1] % 16 == 0:
        with contextlib.redirect_stdout(io.StringIO()):
            dec = rcodec.jpeg_decode(hi)
        assert np.array_equal(rcomp.jpeg_decompression(dec), orc.jpeg_decompression(planes))


WAVELET_SETTINGS = [
    dict(levels=1, multiplier=1, threshold=5, quality_factor=1),
    dict(levels=2, multiplier=2, threshold=3, quality_factor=1),
    dict(levels=4, multiplier=1, threshold=0, quality_factor=0.5),
    dict(levels=5, multiplier=0.5, threshold=5, quality_factor=0.9),
    dict(levels=3, multiplier=1.5, threshold=2.5, quality_factor=0.25),
]


@pytest.mark.reference
@pytest.mark.parametrize("cfg", WAVELET_SETTINGS)
@pytest.mark.parametrize("shape", [(64, 96), (50, 38)])
def test_oracle_wavelet_settings_against_live_reference(cfg, shape):
    """The oracle at non-default wavelet settings (reference settings.py:12-16) against the unmodified
    reference's wavelet_compression / wavelet_decompression (both on the pywt stand-in)."""
    from oracle import refshim
    rsettings = refshim.install()
    import hiccup.compression as rcomp
    import hiccup.model as rmodel
    saved = (rsettings.WAVELET, rsettings.WAVELET_NUM_LEVELS, rsettings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
             rsettings.WAVELET_THRESHOLD, rsettings.WAVELET_QUALITY_FACTOR)
    try:
        rsettings.WAVELET = rmodel.Wavelet.HAAR if cfg["levels"] % 2 else rmodel.Wavelet.DAUBECHIE
        rsettings.WAVELET_NUM_LEVELS = cfg["levels"]
        rsettings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER = cfg["multiplier"]
        rsettings.WAVELET_THRESHOLD = cfg["threshold"]
        rsettings.WAVELET_QUALITY_FACTOR = cfg["quality_factor"]
        rgb = orc.synthetic_image(shape[0], shape[1], 300 + cfg["levels"])
        want = rcomp.wavelet_compression(rgb)
        got = orc.wavelet_compression(rgb, **cfg)
        for ch in orc.CHANNELS:
            assert len(got[ch]) == len(want.as_dict[ch]) == 3 * cfg["levels"] + 1
            for a, b in zip(got[ch], want.as_dict[ch]):
                assert a.shape == b.shape and np.array_equal(a, b)
        assert np.array_equal(orc.wavelet_decompression(got, cfg["multiplier"]), rcomp.wavelet_decompression(want))
    finally:
        (rsettings.WAVELET, rsettings.WAVELET_NUM_LEVELS, rsettings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
         rsettings.WAVELET_THRESHOLD, rsettings.WAVELET_QUALITY_FACTOR) = saved


WS_GOLDENS = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "ws_*.npz")))


@pytest.mark.parametrize("name", WS_GOLDENS)
def test_oracle_equals_reference_wavelet_settings_goldens(name):
    """Goldens the UNMODIFIED reference wrote at non-default wavelet settings (tests/golden/gen_golden.py, on the pywt
    stand-in): sub-bands, run-length streams, tables, bit strings and -- where the reference's own decoder reads its
    file back -- decoded pixels."""
    import json
    g = load_golden(name)
    cfg = json.loads(str(g["settings"]))
    kw = {k: cfg[k] for k in ("levels", "multiplier", "threshold", "quality_factor")}
    planes = orc.wavelet_compression(g["rgb"], **kw)
    n_bands = 3 * cfg["levels"] + 1
    for ch in orc.CHANNELS:
        assert len(planes[ch]) == n_bands
        for i in range(n_bands):
            assert np.array_equal(planes[ch][i], g["band_%s_%d" % (ch, i)])
    st = orc.wavelet_streams(planes)
    for ch in orc.CHANNELS:
        assert np.array_equal(st["value"][ch], g["rle_val_" + ch])
        assert np.array_equal(st["length"][ch], g["rle_len_" + ch])
    enc = orc.wavelet_encode(planes)
    tables, bits, shapes = _stream_content(pickle.loads(g["hic"].tobytes()), 6)
    assert [[(int(a), b) for a, b in t] for t in enc["tables"]] == tables
    assert [orc.padded_bits_to_bytes(b) for b in enc["bits"]] == bits
    assert enc["shapes"] == shapes
    if str(g["decode_error"]) == "":
        dec = orc.wavelet_decode(enc)
        for ch in orc.CHANNELS:
            for i in range(n_bands):
                assert np.array_equal(dec[ch][i], planes[ch][i])
        assert np.array_equal(orc.wavelet_decompression(dec, cfg["multiplier"]), g["rgb_out"])
