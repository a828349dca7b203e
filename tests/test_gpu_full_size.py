"""BASELINE.json's configurations at their full sizes, through size-independent properties and through
the oracle on the largest sample it finishes in seconds (SURVEY 8(c) "large sizes")."""
import hashlib

import numpy as np
import pytest

from oracle import hiccup_oracle as orc

pytestmark = pytest.mark.gpu


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _batch(n, h, w, distinct, seed):
    base = np.stack([orc.synthetic_image(h, w, seed + i) for i in range(distinct)])
    return np.concatenate([base] * (-(-n // distinct)))[:n]


def _oracle_streams(rgb):
    planes = orc.jpeg_compression(rgb)
    enc = orc.jpeg_encode(planes)
    return planes, enc


def _check_image_against_oracle(codec, enc, rgb, i, out):
    """Image i of a batch: tables, framed bit strings and decoded pixels equal the oracle's."""
    planes, want = _oracle_streams(rgb)
    for kind in range(3):
        for c in range(3):
            s = (i * 3 + c) * 3 + kind
            k = kind * 3 + c
            assert enc.table(s) == [(int(a), b) for a, b in want["tables"][k]], "image %d table %d" % (i, k)
            assert enc.framed(s) == orc.padded_bits_to_bytes(want["bits"][k]), "image %d bits %d" % (i, k)
    assert np.array_equal(out[i], orc.jpeg_decompression(planes)), "image %d pixels" % i


def test_c2_batch_of_1024_640x426():
    """configs[1]: every image of the batch is independent, so (a) images that repeat in the input must give
    identical streams and pixels wherever they sit in the batch, (b) the device-resident round trip equals
    the host-to-host one, (c) samples equal the oracle bit for bit."""
    from hiccup_b200.batch import DctBatchCodec
    n, h, w, distinct = 1024, 426, 640, 8
    rgb = _batch(n, h, w, distinct, 4000)
    codec = DctBatchCodec(n, h, w)
    enc = codec.encode(rgb)
    out = codec.decode(enc).copy()
    assert int(codec.forward_stats[3]) == 0 and int(codec.inverse_stats[3]) == 0
    for i in (0, 3):
        _check_image_against_oracle(codec, enc, rgb[i], i, out)
    for i in range(distinct, n, 97):                          # a repeat equals its first occurrence
        j = i % distinct
        for k in range(9):
            assert enc.framed(i * 9 + k) == enc.framed(j * 9 + k)
            assert enc.table(i * 9 + k) == enc.table(j * 9 + k)
        assert _digest(out[i]) == _digest(out[j])
    codec.upload(rgb)
    codec.encode_device()
    codec.decode_device()
    dev = codec.d_out.download(np.uint8, out.size).reshape(out.shape)
    assert _digest(dev) == _digest(out)
    codec.close()


def test_c3_batch_of_4k_images():
    """configs[2] at its full size (256 x 3840x2160): repeats are identical wherever they sit in the batch, one
    whole 4K image equals the oracle (streams and pixels)."""
    from hiccup_b200.batch import DctBatchCodec
    n, h, w, distinct = 256, 2160, 3840, 2
    rgb = _batch(n, h, w, distinct, 4100)
    codec = DctBatchCodec(n, h, w)
    enc = codec.encode(rgb)
    out = codec.decode(enc).copy()
    _check_image_against_oracle(codec, enc, rgb[1], 1, out)
    for i in (2, 17, 63, 128, 255):
        j = i % distinct
        assert all(enc.framed(i * 9 + k) == enc.framed(j * 9 + k) for k in range(9))
        assert _digest(out[i]) == _digest(out[j])
    codec.close()


def test_c4_wavelet_8k():
    """configs[3]: 7680x4320 wavelet mode.  Whole-image comparison with the oracle (sub-bands through the
    drop-in API, streams and decoded pixels through the batched one)."""
    from hiccup_b200.batch import WaveletBatchCodec
    h, w = 4320, 7680
    rgb = np.tile(orc.synthetic_image(1080, 1920, 4200), (4, 4, 1))
    planes = orc.wavelet_compression(rgb)
    want = orc.wavelet_encode(planes)
    codec = WaveletBatchCodec(1, h, w)
    enc = codec.encode(rgb[None])
    for kind in (1, 2):
        for c in range(3):
            k = (kind - 1) * 3 + c
            assert enc.table(c * 3 + kind) == [(int(a), b) for a, b in want["tables"][k]], "table %d" % k
            assert enc.framed(c * 3 + kind) == orc.padded_bits_to_bytes(want["bits"][k]), "bits %d" % k
    out = codec.decode(enc)
    assert np.array_equal(out[0], orc.wavelet_decompression(orc.wavelet_decode(want)))
    codec.close()


def test_c5_16384_square_in_eight_bands():
    """configs[4] on one GPU: eight row bands stitched on the host equal the one-band encode byte for
    byte, and a 512x512 window of coefficients equals the oracle's transform of that window (away from the
    window's own border, where the chroma pyramid sees different neighbours)."""
    from hiccup_b200 import _lib, bands
    from hiccup_b200.batch import DctBatchCodec
    size = 16384
    rgb = np.tile(orc.synthetic_image(2048, 2048, 4300), (8, 8, 1))
    eight = bands.encode_banded(rgb, 8).byte_stream()
    one = bands.encode_banded(rgb, 1).byte_stream()
    assert len(eight) == len(one) == 21
    for i, (a, b) in enumerate(zip(eight, one)):
        assert a == b, "payload %d" % i
    codec = DctBatchCodec(1, size, size)
    codec.upload(rgb[None])
    codec._forward()
    coef = codec.coefficients()[0]
    g = codec.g
    y0, x0, win = 7168, 9216, 512
    want = orc.jpeg_compression(rgb[y0:y0 + win, x0:x0 + win])
    zz = orc.blocks_zigzag(want["lum"]).reshape(win // 8, win // 8, 64)
    lum = coef[:g.nb_l].reshape(g.nby_l, g.nbx_l, 64)[y0 // 8:(y0 + win) // 8, x0 // 8:(x0 + win) // 8]
    assert np.array_equal(lum, zz)
    zc = orc.blocks_zigzag(want["cr"]).reshape(win // 16, win // 16, 64)
    cr = coef[g.nb_l:g.nb_l + g.nb_c].reshape(g.nby_c, g.nbx_c, 64)[y0 // 16:(y0 + win) // 16, x0 // 16:(x0 + win) // 16]
    assert np.array_equal(cr[1:-1, 1:-1], zc[1:-1, 1:-1])
    # decode of the stitched stream: the same window of PIXELS against the oracle's decode of the window's own
    # coefficients (equal to the image's there, as just checked), away from the window border where the chroma
    # pyramid (pyrUp's 5-tap filter on top of the border blocks above) sees other neighbours
    from hiccup_b200 import codec as codec_mod, compression
    hic = bands.encode_banded(rgb, 8)
    out = compression.jpeg_decompression(codec_mod.jpeg_decode(hic))
    assert out.shape == rgb.shape
    ref = orc.jpeg_decompression(want)
    m = 32
    assert np.array_equal(out[y0 + m:y0 + win - m, x0 + m:x0 + win - m], ref[m:-m, m:-m])
    # and the whole image through a size-independent property: the input repeats with period 2048, so must the
    # output away from the image border
    assert _digest(out[2048:4096, 2048:4096]) == _digest(out[4096:6144, 6144:8192])
    codec.close()
