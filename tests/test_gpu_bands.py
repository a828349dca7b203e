"""Row-band sharded encode on the GPU (hiccup_b200/bands.py): K bands of one image, each through K1 and
the two-pass entropy kernels with the seam state, stitched on the host -- must be byte-identical to the
whole-image encode (itself pinned to the reference)."""
import numpy as np
import pytest

from oracle import hiccup_oracle as orc

pytestmark = pytest.mark.gpu


def _image(kind, h, w, seed):
    if kind == "synthetic":
        return orc.synthetic_image(h, w, seed)
    img = np.full((h, w, 3), 90, np.uint8)
    if kind == "half":
        img[:h // 2] = orc.synthetic_image(h // 2, w, seed)
    elif kind == "bottom":
        img[3 * h // 4:] = orc.synthetic_image(h - 3 * h // 4, w, seed)
    elif kind == "middle":
        img[h // 3:h // 3 + 24] = orc.synthetic_image(24, w, seed)
    return img


@pytest.mark.parametrize("kind,shape,k", [("synthetic", (96, 80), 2), ("synthetic", (130, 72), 3), ("synthetic", (426, 640), 4),
                                          ("flat", (64, 48), 3), ("half", (128, 64), 4), ("bottom", (256, 64), 8),
                                          ("middle", (192, 112), 6), ("synthetic", (50, 34), 2),
                                          ("synthetic", (1080, 1920), 8), ("synthetic", (2048, 512), 5)])
def test_banded_encode_equals_whole_image(kind, shape, k):
    from hiccup_b200 import bands, codec, compression
    rgb = _image(kind, shape[0], shape[1], 17)
    want = codec.jpeg_encode(compression.jpeg_compression(rgb)).byte_stream()
    got = bands.encode_banded(rgb, k).byte_stream()
    assert len(got) == len(want) == 21
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, "payload %d differs (%d vs %d bytes)" % (i, len(a), len(b))


def test_banded_encode_equals_oracle_and_decodes():
    from hiccup_b200 import bands, codec, compression, hicimage
    rgb = orc.synthetic_image(160, 96, 23)
    planes = orc.jpeg_compression(rgb)
    enc = orc.jpeg_encode(planes)
    hi = bands.encode_banded(rgb, 3)
    stream = hi.byte_stream()
    for i in range(9):
        assert [(int(a), b) for a, b in hi.payloads[i].rows] == [(int(a), b) for a, b in enc["tables"][i]]
        assert stream[10 + i] == orc.padded_bits_to_bytes(enc["bits"][i])
    out = compression.jpeg_decompression(codec.jpeg_decode(hicimage.HicImage.from_bytes(stream)))
    assert np.array_equal(out, orc.jpeg_decompression(planes))


@pytest.mark.parametrize("shape,k", [((426, 640), 4), ((1080, 1920), 8), ((96, 80), 2)])
def test_device_side_stitching_equals_host_stitching(shape, k):
    """bands.upload_stitched (band strings copied to their places on the device) against bands.stitch_into, and the
    decode of what it built against the decode of the host-stitched file."""
    from hiccup_b200 import _lib, bands
    from hiccup_b200.batch import DctBatchCodec
    rgb = orc.synthetic_image(shape[0], shape[1], 31)
    cuts = bands.plan_bands(shape[0], k)
    workers = []
    for b in range(len(cuts) - 1):
        wk = bands.BandWorker(b, len(cuts) - 1, shape[0], shape[1], cuts[b], cuts[b + 1])
        wk.load(rgb)
        workers.append(wk)
    try:
        res = bands.run_local(workers)[0]
        size = bands.stitch_layout(res["all_bits"], 9)[3]
        host = np.zeros(size, np.uint8)
        off, length, nbits = bands.stitch_into(host, res["band_bytes"], res["all_bits"], 9)
        d = _lib.DeviceBuffer(size)
        off_d, nbits_d = bands.upload_stitched(res, d)
        assert off_d.tolist() == list(off) and nbits_d.tolist() == list(nbits)
        assert np.array_equal(d.download(np.uint8, size), host)
        codec = DctBatchCodec(1, shape[0], shape[1])
        want = codec.decode(bands.to_encoded_streams(res, shape[0], shape[1])).copy()
        index, syms, packed = res["tables"]
        codec.decoder.decode_device_data(index, syms, packed, d.ptr, off_d, nbits_d, codec.d_coef_dec.ptr)
        codec._inverse()
        assert np.array_equal(codec.fetch(), want)
        assert np.array_equal(want[0], orc.jpeg_decompression(orc.jpeg_compression(rgb)))
        codec.close()
        d.free()
    finally:
        for wk in workers:
            wk.close()


@pytest.mark.parametrize("shape,k", [((426, 640), 4), ((2048, 512), 5), ((96, 80), 2), ((1080, 1920), 8)])
def test_sharded_decode_bands_equal_whole_image_decode(shape, k):
    """bands.BandDecoder: every band runs K7 / K8 on its own block rows (+ halo) of the image's coefficients; its rows
    must be the whole-image decode's rows (and the oracle's), seams included."""
    from hiccup_b200 import bands
    from hiccup_b200.batch import DctBatchCodec
    h, w = shape
    rgb = orc.synthetic_image(h, w, 41)
    codec = DctBatchCodec(1, h, w)
    codec.upload(rgb[None])
    codec._forward()
    coef = codec.coefficients()[0]
    codec.d_coef_dec.upload(coef, codec.stream)
    codec._inverse()
    want = codec.fetch().reshape(codec.out_h, codec.out_w, 3).copy()
    assert np.array_equal(want, orc.jpeg_decompression(orc.jpeg_compression(rgb)))
    cuts = bands.plan_bands(h, k)
    got = np.zeros_like(want)
    for b in range(len(cuts) - 1):
        dec = bands.BandDecoder(h, w, cuts[b], cuts[b + 1])
        try:
            mine = np.concatenate([coef[a:z] for a, z in dec.src])
            assert mine.shape[0] == dec.g.blocks_per_image
            dec.codec.d_coef_dec.upload(mine, dec.codec.stream)
            dec.inverse()
            rows = np.empty((cuts[b + 1] - cuts[b], codec.out_w, 3), np.uint8)
            dec.fetch_rows(rows)
            got[cuts[b]:cuts[b + 1]] = rows
        finally:
            dec.close()
    assert np.array_equal(got, want)
    codec.close()

