"""GPU parity of the wavelet mode at settings other than the defaults (reference settings.py:12-16):
WAVELET_NUM_LEVELS 1..5, the haar alias, multiplier, threshold, and a quality factor below 1 (the order
statistic of quantization.py:84-94 taken on the device).  Target: the oracle, which tests/test_oracle_pinning.py
pins against the unmodified reference at the same settings (both on the pywt stand-in: PARITY UNPINNED against
the real PyWavelets)."""
import contextlib

import numpy as np
import pytest

from oracle import hiccup_oracle as orc

pytestmark = pytest.mark.gpu

CH = ("lum", "cr", "cb")
SETTINGS = [
    dict(levels=1, multiplier=1, threshold=5, quality_factor=1),
    dict(levels=2, multiplier=2, threshold=3, quality_factor=1),
    dict(levels=3, multiplier=1.5, threshold=2.5, quality_factor=0.25),
    dict(levels=3, multiplier=1, threshold=5, quality_factor=0.6),
    dict(levels=4, multiplier=1, threshold=0, quality_factor=0.5),
    dict(levels=5, multiplier=0.5, threshold=5, quality_factor=0.9),
    dict(levels=5, multiplier=1, threshold=5, quality_factor=1),
]


@contextlib.contextmanager
def wavelet_settings(cfg, haar=False):
    from hiccup_b200 import model, settings
    saved = (settings.WAVELET, settings.WAVELET_NUM_LEVELS, settings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
             settings.WAVELET_THRESHOLD, settings.WAVELET_QUALITY_FACTOR)
    try:
        settings.WAVELET = model.Wavelet.HAAR if haar else model.Wavelet.DAUBECHIE
        settings.WAVELET_NUM_LEVELS = cfg["levels"]
        settings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER = cfg["multiplier"]
        settings.WAVELET_THRESHOLD = cfg["threshold"]
        settings.WAVELET_QUALITY_FACTOR = cfg["quality_factor"]
        yield
    finally:
        (settings.WAVELET, settings.WAVELET_NUM_LEVELS, settings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
         settings.WAVELET_THRESHOLD, settings.WAVELET_QUALITY_FACTOR) = saved


def _same(a, b):
    return all(len(a[ch]) == len(b[ch]) and
               all(np.asarray(x).shape == np.asarray(y).shape and np.array_equal(x, y) for x, y in zip(a[ch], b[ch]))
               for ch in CH)


@pytest.mark.parametrize("cfg", SETTINGS)
@pytest.mark.parametrize("shape,seed", [((64, 96), 1), ((50, 38), 2), ((131, 70), 3), ((256, 320), 4), ((7, 5), 5)])
def test_compression_and_decompression_match_oracle(cfg, shape, seed):
    """wavelet_compression and wavelet_decompression (without the codec in between: odd shapes go through
    waverec2's trim rule) at every setting."""
    from hiccup_b200 import compression, model
    rgb = orc.synthetic_image(shape[0], shape[1], 700 + seed)
    with wavelet_settings(cfg, haar=bool(seed & 1)):
        got = compression.wavelet_compression(rgb)
        want = orc.wavelet_compression(rgb, **cfg)
        assert all(b.dtype == np.int32 for ch in CH for b in got.as_dict[ch])
        assert _same(got.as_dict, want), "sub-bands differ at %r" % (cfg,)
        out = compression.wavelet_decompression(model.CompressedImage.from_dict(want))
        ref = orc.wavelet_decompression(want, cfg["multiplier"])
        assert out.shape == ref.shape and np.array_equal(out, ref), "%d pixels differ" % int((out != ref).sum())


@pytest.mark.parametrize("cfg", [c for c in SETTINGS if c["levels"] in (2, 3, 5)])
def test_file_round_trip_matches_oracle(cfg):
    """pixels -> .hic payloads -> pixels at the level counts the reference's decoder reads correctly (2, 3, 5:
    codec.py:182-189), every stage against the oracle."""
    from hiccup_b200 import codec, compression, hicimage
    rgb = orc.synthetic_image(128, 192, 900 + cfg["levels"])
    with wavelet_settings(cfg):
        planes = orc.wavelet_compression(rgb, **cfg)
        enc = orc.wavelet_encode(planes)
        hi = codec.wavelet_encode(compression.wavelet_compression(rgb))
        stream = hi.byte_stream()
        for i in range(6):
            assert [(int(a), b) for a, b in hi.payloads[i].rows] == [(int(a), b) for a, b in enc["tables"][i]], "table %d" % i
            assert stream[7 + i] == orc.padded_bits_to_bytes(enc["bits"][i]), "bit string %d" % i
        assert tuple(hi.payloads[12].numbers) == enc["shapes"][0] and tuple(hi.payloads[13].numbers) == enc["shapes"][1]
        back = codec.wavelet_decode(hicimage.HicImage.from_bytes(stream))
        assert _same(back.as_dict, planes)
        out = compression.wavelet_decompression(back)
        want = orc.wavelet_decompression(orc.wavelet_decode(enc), cfg["multiplier"])
        assert np.array_equal(out, want)


@pytest.mark.parametrize("levels", [1, 4])
def test_decode_refuses_level_counts_the_reference_misreads(levels):
    from hiccup_b200 import codec, compression
    cfg = dict(levels=levels, multiplier=1, threshold=5, quality_factor=1)
    with wavelet_settings(cfg):
        hi = codec.wavelet_encode(compression.wavelet_compression(orc.synthetic_image(64, 64, 5)))
        with pytest.raises(ValueError):
            codec.wavelet_decode(hi)


def test_batch_codec_at_general_settings():
    """WaveletBatchCodec on the general kernels: encode + decode of a small batch equals the oracle per image."""
    from hiccup_b200.batch import WaveletBatchCodec
    cfg = dict(levels=2, multiplier=2, threshold=3, quality_factor=0.7)
    n, h, w = 3, 64, 96
    rgb = np.stack([orc.synthetic_image(h, w, 40 + i) for i in range(n)])
    with wavelet_settings(cfg):
        codec = WaveletBatchCodec(n, h, w)
        enc = codec.encode(rgb)
        out = codec.decode(enc).reshape(n, h, w, 3).copy()
        for i in range(n):
            planes = orc.wavelet_compression(rgb[i], **cfg)
            want = orc.wavelet_encode(planes)
            for kind in range(2):
                for c in range(3):
                    s = (i * 3 + c) * 3 + kind + 1
                    assert enc.framed(s) == orc.padded_bits_to_bytes(want["bits"][kind * 3 + c])
            assert np.array_equal(out[i], orc.wavelet_decompression(planes, cfg["multiplier"]))
        codec.close()


def test_unsupported_settings_raise():
    from hiccup_b200 import compression, model, settings
    rgb = orc.synthetic_image(32, 32, 1)
    saved = settings.WAVELET
    try:
        settings.WAVELET = model.Wavelet.COIF
        with pytest.raises(NotImplementedError):
            compression.wavelet_compression(rgb)
    finally:
        settings.WAVELET = saved
    with wavelet_settings(dict(levels=3, multiplier=0, threshold=5, quality_factor=1)):
        with pytest.raises(NotImplementedError):
            compression.wavelet_compression(rgb)
    with wavelet_settings(dict(levels=6, multiplier=1, threshold=5, quality_factor=1)):
        with pytest.raises(NotImplementedError):
            compression.wavelet_compression(rgb)


import glob
import json
import os
import pickle

from tests.conftest import GOLDEN, load_golden

WS_GOLDENS = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "ws_*.npz")))


@pytest.mark.parametrize("name", WS_GOLDENS)
def test_reference_goldens_at_general_settings(name):
    """What the UNMODIFIED reference wrote at non-default settings (on the pywt stand-in): sub-bands from pixels, the
    `.hic` payloads byte for byte, the reference's file decoded, and the pixels it decompressed to."""
    from hiccup_b200 import codec, compression, hicimage, model
    g = load_golden(name)
    cfg = json.loads(str(g["settings"]))
    n_bands = 3 * cfg["levels"] + 1
    want_hic = pickle.loads(g["hic"].tobytes())
    with wavelet_settings(cfg, haar=cfg["haar"]):
        got = compression.wavelet_compression(g["rgb"])
        for ch in CH:
            assert len(got.as_dict[ch]) == n_bands
            for i in range(n_bands):
                assert np.array_equal(got.as_dict[ch][i], g["band_%s_%d" % (ch, i)]), "%s band %d" % (ch, i)
        assert codec.wavelet_encode(got).byte_stream() == want_hic
        if str(g["decode_error"]) == "":
            back = codec.wavelet_decode(hicimage.HicImage.from_bytes(want_hic))
            assert _same(back.as_dict, got.as_dict)
            assert np.array_equal(compression.wavelet_decompression(back), g["rgb_out"])
        else:           # a level count the reference's own decoder mis-reads: refused here
            with pytest.raises(ValueError):
                codec.wavelet_decode(hicimage.HicImage.from_bytes(want_hic))
            golden = model.CompressedImage.from_dict({ch: [g["band_%s_%d" % (ch, i)] for i in range(n_bands)] for ch in CH})
            out = compression.wavelet_decompression(golden)
            kw = {k: cfg[k] for k in ("levels", "multiplier", "threshold", "quality_factor")}
            assert np.array_equal(out, orc.wavelet_decompression(orc.wavelet_compression(g["rgb"], **kw), cfg["multiplier"]))
