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
JPEG = [n for n in ALL if not n.startswith("w_")]
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
