"""GPU parity of the wavelet ("HIC") mode (K9, K10, flat-mode entropy stage).

Targets: (a) goldens recorded from the UNMODIFIED reference running on oracle/pywt_standin.py
(tests/golden/w_*.npz: sub-band values, `.hic` bytes, decoded pixels), (b) the oracle on more shapes,
including ones larger than a CTA region and shapes that need PyWavelets' symmetric extension.
PyWavelets itself is not installable here, so this mode is parity-unpinned against the real package."""
import glob
import os
import pickle

import numpy as np
import pytest

from oracle import hiccup_oracle as orc
from tests.conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

W_GOLDENS = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "w_*.npz")))
CH = ("lum", "cr", "cb")


def _golden_comp(g):
    from hiccup_b200 import model
    return model.CompressedImage.from_dict({ch: [g["band_%s_%d" % (ch, i)] for i in range(10)] for ch in CH})


def _same_bands(a, b):
    return all(len(a[ch]) == len(b[ch]) and
               all(np.asarray(x).shape == np.asarray(y).shape and np.array_equal(x, y) for x, y in zip(a[ch], b[ch]))
               for ch in CH)


@pytest.mark.parametrize("name", W_GOLDENS)
def test_wavelet_compression_matches_reference_golden(name):
    from hiccup_b200 import compression
    g = load_golden(name)
    got = compression.wavelet_compression(g["rgb"]).as_dict
    for ch in CH:
        for i in range(10):
            want = g["band_%s_%d" % (ch, i)]
            assert got[ch][i].dtype == np.int32 and got[ch][i].shape == want.shape
            assert np.array_equal(got[ch][i], want), "%s %s band %d" % (name, ch, i)


@pytest.mark.parametrize("name", W_GOLDENS)
def test_wavelet_encode_bytes_equal_reference(name):
    from hiccup_b200 import codec, compression
    g = load_golden(name)
    want = pickle.loads(g["hic"].tobytes())
    got = codec.wavelet_encode(_golden_comp(g)).byte_stream()
    assert len(got) == len(want) == 15
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, "%s: payload %d differs (%d vs %d bytes)" % (name, i, len(a), len(b))
    # and the whole path from pixels
    assert codec.wavelet_encode(compression.wavelet_compression(g["rgb"])).byte_stream() == want


@pytest.mark.parametrize("name", W_GOLDENS)
def test_wavelet_decode_of_reference_file(name):
    from hiccup_b200 import codec, compression, hicimage
    g = load_golden(name)
    hi = hicimage.HicImage.from_bytes(pickle.loads(g["hic"].tobytes()))
    dec = codec.wavelet_decode(hi)
    assert _same_bands(dec.as_dict, _golden_comp(g).as_dict)
    assert all(b.dtype == np.float64 for ch in CH for b in dec.as_dict[ch])
    if str(g["decode_error"]) == "":
        out = compression.wavelet_decompression(dec)
        assert out.dtype == np.uint8 and out.shape == g["rgb_out"].shape
        assert np.array_equal(out, g["rgb_out"]), "%s: %d pixels differ" % (name, int((out != g["rgb_out"]).sum()))


@pytest.mark.parametrize("shape,seed", [((128, 128), 81), ((136, 264), 82), ((256, 384), 83), ((33, 47), 84),
                                        ((50, 50), 85), ((8, 8), 86), ((1, 1), 87), ((131, 70), 88), ((720, 1280), 89)])
def test_wavelet_forward_matches_oracle(shape, seed):
    """K9 against the oracle, including odd shapes (symmetric extension at every level)."""
    from hiccup_b200 import compression
    rgb = orc.synthetic_image(shape[0], shape[1], seed) if min(shape) >= 2 else np.full(shape + (3,), 77, np.uint8)
    got = compression.wavelet_compression(rgb).as_dict
    want = orc.wavelet_compression(rgb)
    for ch in CH:
        for i in range(10):
            assert got[ch][i].shape == want[ch][i].shape, "%s band %d shape" % (ch, i)
            assert np.array_equal(got[ch][i], want[ch][i]), "%s band %d: %d differ" % (
                ch, i, int((got[ch][i] != want[ch][i]).sum()))


@pytest.mark.parametrize("shape,seed", [((128, 128), 91), ((136, 264), 92), ((256, 384), 93), ((720, 1280), 94)])
def test_wavelet_round_trip_matches_oracle(shape, seed):
    """rgb -> compression -> encode -> bytes -> decode -> decompression, every stage against the oracle."""
    from hiccup_b200 import codec, compression, hicimage
    rgb = orc.synthetic_image(shape[0], shape[1], seed)
    planes = orc.wavelet_compression(rgb)
    enc = orc.wavelet_encode(planes)
    hi = codec.wavelet_encode(compression.wavelet_compression(rgb))
    stream = hi.byte_stream()
    for i in range(6):
        assert [(int(a), b) for a, b in hi.payloads[i].rows] == [(int(a), b) for a, b in enc["tables"][i]], "table %d" % i
        assert stream[7 + i] == orc.padded_bits_to_bytes(enc["bits"][i]), "bit string %d" % i
    assert tuple(hi.payloads[12].numbers) == enc["shapes"][0] and tuple(hi.payloads[13].numbers) == enc["shapes"][1]
    back = codec.wavelet_decode(hicimage.HicImage.from_bytes(stream))
    assert _same_bands(back.as_dict, planes)
    out = compression.wavelet_decompression(back)
    want = orc.wavelet_decompression(orc.wavelet_decode(enc))
    assert np.array_equal(out, want), "%d pixels differ" % int((out != want).sum())


def test_wavelet_noise_and_extremes():
    from hiccup_b200 import compression
    rng = np.random.default_rng(5)
    for rgb in (rng.integers(0, 256, (64, 96, 3), dtype=np.uint8), np.zeros((40, 40, 3), np.uint8),
                np.full((40, 40, 3), 255, np.uint8)):
        got = compression.wavelet_compression(rgb).as_dict
        want = orc.wavelet_compression(rgb)
        assert _same_bands(got, want)


def test_wavelet_decode_rejects_shapes_the_reference_cannot_read():
    from hiccup_b200 import codec, compression
    rgb = orc.synthetic_image(33, 47, 3)
    hi = codec.wavelet_encode(compression.wavelet_compression(rgb))     # encoding works for any shape
    with pytest.raises(ValueError):
        codec.wavelet_decode(hi)


def test_long_huffman_codes_round_trip():
    """Fibonacci symbol frequencies give codes far longer than the decoder's 20-bit table levels (the
    predecessor-search path of D1); the coded bits must still equal the oracle's and decode back."""
    from hiccup_b200 import codec, hicimage, model
    h = w = 512
    shapes = [(64, 64)] * 4 + [(128, 128)] * 3 + [(256, 256)] * 3
    total = sum(a * b for a, b in shapes)
    fib = [1, 1]
    while sum(fib) + fib[-1] + fib[-2] <= total:
        fib.append(fib[-1] + fib[-2])
    fib[-1] += total - sum(fib)              # no zeros at all: nothing but the Fibonacci chain in the tree
    rng = np.random.default_rng(11)
    planes = {}
    for ci, ch in enumerate(CH):
        vals = np.concatenate([np.full(f, 7 + k + ci, np.int32) for k, f in enumerate(fib)])
        stream = rng.permutation(vals)
        bands, off = [], 0
        for (bh, bw) in shapes:
            bands.append(stream[off:off + bh * bw].reshape(bh, bw).copy())
            off += bh * bw
        planes[ch] = bands
    enc = orc.wavelet_encode(planes)
    assert max(len(c) for _, c in enc["tables"][0]) > 22, "the test must produce codes past the table levels"
    hi = codec.wavelet_encode(model.CompressedImage.from_dict(planes))
    stream_bytes = hi.byte_stream()
    for i in range(6):
        assert [(int(a), b) for a, b in hi.payloads[i].rows] == [(int(a), b) for a, b in enc["tables"][i]], "table %d" % i
        assert stream_bytes[7 + i] == orc.padded_bits_to_bytes(enc["bits"][i]), "bit string %d" % i
    back = codec.wavelet_decode(hicimage.HicImage.from_bytes(stream_bytes))
    assert _same_bands(back.as_dict, planes)


def test_codes_that_never_self_synchronise_round_trip():
    """Eight equally frequent values give a complete tree of 3-bit codes: codewords straddle the decoder's
    128-bit subsequences in a fixed phase, so a wrong start never finds the true boundaries by itself and the
    resynchronisation needs one round per 4 KB tile (24 here).  The decode must still be exact."""
    from hiccup_b200 import codec, hicimage, model
    shapes = [(64, 64)] * 4 + [(128, 128)] * 3 + [(256, 256)] * 3
    total = sum(a * b for a, b in shapes)
    rng = np.random.default_rng(12)
    planes = {}
    for ci, ch in enumerate(CH):
        vals = np.repeat(np.arange(1, 9, dtype=np.int32) * (ci + 1), total // 8)
        stream = rng.permutation(vals)
        bands, off = [], 0
        for (bh, bw) in shapes:
            bands.append(stream[off:off + bh * bw].reshape(bh, bw).copy())
            off += bh * bw
        planes[ch] = bands
    enc = orc.wavelet_encode(planes)
    assert set(len(c) for _, c in enc["tables"][0]) == {3}
    hi = codec.wavelet_encode(model.CompressedImage.from_dict(planes))
    stream_bytes = hi.byte_stream()
    for i in range(6):
        assert stream_bytes[7 + i] == orc.padded_bits_to_bytes(enc["bits"][i]), "bit string %d" % i
    back = codec.wavelet_decode(hicimage.HicImage.from_bytes(stream_bytes))
    assert _same_bands(back.as_dict, planes)
    # and with restart records there is nothing to synchronise
    back2 = codec.wavelet_decode(codec.add_restart_records(hi))
    assert _same_bands(back2.as_dict, planes)
