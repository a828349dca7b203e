"""GPU parity of the entropy stage (E1-E3, D1-D3) against the reference goldens and the oracle."""
import glob
import os
import pickle

import numpy as np
import pytest

from oracle import hiccup_oracle as orc
from tests.conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

JPEG_GOLDENS = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz"))
                      if not os.path.basename(f).startswith(("w_", "ws_")))


def _golden_comp(g):
    from hiccup_b200 import model
    return model.CompressedImage(g["coef_lum"], g["coef_cr"], g["coef_cb"])


@pytest.mark.parametrize("name", JPEG_GOLDENS)
def test_jpeg_encode_bytes_equal_reference(name):
    """codec.jpeg_encode(...).byte_stream() is byte-identical to the reference's .hic payload list."""
    from hiccup_b200 import codec
    g = load_golden(name)
    want = pickle.loads(g["hic"].tobytes())
    got = codec.jpeg_encode(_golden_comp(g)).byte_stream()
    assert len(got) == len(want) == 21
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, "%s: payload %d differs (%d vs %d bytes)" % (name, i, len(a), len(b))


@pytest.mark.parametrize("name", JPEG_GOLDENS)
def test_full_encode_pipeline_bytes_equal_reference(name):
    """rgb -> jpeg_compression -> jpeg_encode -> bytes: the whole encode path, bit-exact .hic."""
    from hiccup_b200 import codec, compression
    g = load_golden(name)
    want = pickle.loads(g["hic"].tobytes())
    got = codec.jpeg_encode(compression.jpeg_compression(g["rgb"])).byte_stream()
    assert got == want


@pytest.mark.parametrize("name", JPEG_GOLDENS)
def test_jpeg_decode_of_reference_file(name):
    from hiccup_b200 import codec, hicimage
    g = load_golden(name)
    hi = hicimage.HicImage.from_bytes(pickle.loads(g["hic"].tobytes()))
    dec = codec.jpeg_decode(hi)
    for ch, arr in dec.as_dict.items():
        assert arr.dtype == np.float64
        assert np.array_equal(arr, g["coef_" + ch]), "%s %s" % (name, ch)


@pytest.mark.parametrize("shape,seed", [((426, 640), 41), ((1080, 1920), 42), ((40, 24), 43), ((136, 264), 44)])
def test_symbol_streams_match_oracle(shape, seed):
    from hiccup_b200 import _lib, compression, entropy, model
    rgb = orc.synthetic_image(shape[0], shape[1], seed)
    planes = orc.jpeg_compression(rgb)
    want = orc.jpeg_streams(planes)
    coef, g = compression.planes_to_device_blocks(model.CompressedImage(planes["lum"], planes["cr"], planes["cb"]))
    layout = _lib.layout_dct(1, shape[0], shape[1])
    enc = entropy.EntropyEncoder(layout)
    res = enc.encode(coef.ptr)
    dc, val, ln = enc.symbol_arrays()
    for c, ch in enumerate(orc.CHANNELS):
        b0 = int(layout.block_off[c])
        nb = int(layout.nb[c])
        assert np.array_equal(dc[b0:b0 + nb], want["dc"][ch])
        n = int(res.nsym[c * 3 + 1])
        assert n == len(want["ac_value"][ch]) == int(res.nsym[c * 3 + 2])
        assert np.array_equal(val[b0 * 64:b0 * 64 + n], want["ac_value"][ch])
        assert np.array_equal(ln[b0 * 64:b0 * 64 + n], want["ac_length"][ch])
    # tables and bits against the oracle's Huffman stage
    enc_o = orc.jpeg_encode(planes)
    for kind in range(3):
        for c in range(3):
            i = kind * 3 + c
            s = c * 3 + kind
            assert res.table(s) == [(int(a), b) for a, b in enc_o["tables"][i]]
            assert res.framed(s) == orc.padded_bits_to_bytes(enc_o["bits"][i])
    enc.close()
    coef.free()


def test_rle_edge_cases():
    """All-zero planes, a single trailing non-zero, runs of exactly 15/16/30 zeros across blocks."""
    from hiccup_b200 import codec, model
    rng = np.random.default_rng(3)
    cases = []
    z = np.zeros((32, 32), np.int32)
    cases.append((z, z[:16, :16], z[:16, :16]))
    a = z.copy(); a[31, 31] = 7                       # last scan position of the last block
    cases.append((a, z[:16, :16], z[:16, :16]))
    b = z.copy(); b[0, 0] = -5; b[8, 1] = 3           # DC only + one AC far away
    cases.append((b, z[:16, :16], z[:16, :16]))
    c = (rng.integers(-3, 4, (32, 32)) * (rng.random((32, 32)) < 0.05)).astype(np.int32)
    cases.append((c, c[:16, :16].copy(), c[16:, 16:].copy()))
    d = (rng.integers(-300, 300, (64, 48)) * (rng.random((64, 48)) < 0.01)).astype(np.int32)
    cases.append((d, d[:32, :24].copy(), d[32:, 24:].copy()))
    for lum, cr, cb in cases:
        planes = {"lum": lum, "cr": cr, "cb": cb}
        want = orc.jpeg_encode(planes)
        hi = codec.jpeg_encode(model.CompressedImage(lum, cr, cb))
        for i in range(9):
            assert [(int(a_), b_) for a_, b_ in hi.payloads[i].rows] == [(int(a_), b_) for a_, b_ in want["tables"][i]]
            assert hi.payloads[9 + i].byte_stream == orc.padded_bits_to_bytes(want["bits"][i])
        back = codec.jpeg_decode(hi)
        assert np.array_equal(back.luminance_component, lum)
        assert np.array_equal(back.red_chrominance_component, cr)
        assert np.array_equal(back.blue_chrominance_component, cb)


@pytest.mark.parametrize("shape,seed", [((426, 640), 51), ((250, 130), 52), ((1080, 1920), 53)])
def test_round_trip_through_both_stages(shape, seed):
    from hiccup_b200 import codec, compression, hicimage
    rgb = orc.synthetic_image(shape[0], shape[1], seed)
    comp = compression.jpeg_compression(rgb)
    stream = codec.jpeg_encode(comp).byte_stream()
    back = codec.jpeg_decode(hicimage.HicImage.from_bytes(stream))
    assert back == comp
    out = compression.jpeg_decompression(back)
    want = orc.jpeg_decompression(orc.jpeg_compression(rgb))
    assert out.shape == want.shape
    assert np.array_equal(out, want)


@pytest.mark.parametrize("shape,seed,n", [((426, 640), 61, 3), ((64, 64), 62, 5), ((1080, 1920), 63, 1)])
def test_device_huffman_builder_equals_host(shape, seed, n):
    """E2 on the GPU (one CTA per stream replaying heapq) produces the host replay's tables and bits."""
    from hiccup_b200.batch import DctBatchCodec
    rgb = np.stack([orc.synthetic_image(shape[0], shape[1], seed + i) for i in range(n)])
    a = DctBatchCodec(n, shape[0], shape[1], device_codes=True)
    b = DctBatchCodec(n, shape[0], shape[1], device_codes=False)
    ea, eb = a.encode(rgb), b.encode(rgb)
    assert np.array_equal(ea.rows, eb.rows) and np.array_equal(ea.nsym, eb.nsym) and np.array_equal(ea.nbits, eb.nbits)
    for s in range(9 * n):
        assert ea.table(s) == eb.table(s), "stream %d" % s
        assert ea.framed(s) == eb.framed(s), "stream %d" % s
    a.close()
    b.close()


def test_batch_codec_matches_single_image_path():
    from hiccup_b200 import codec, compression
    from hiccup_b200.batch import DctBatchCodec
    n, h, w = 4, 72, 104
    rgb = np.stack([orc.synthetic_image(h, w, 70 + i) for i in range(n)])
    bc = DctBatchCodec(n, h, w)
    enc = bc.encode(rgb)
    images = bc.hic_images(enc)
    out = bc.decode(enc)
    for i in range(n):
        single = codec.jpeg_encode(compression.jpeg_compression(rgb[i]))
        assert images[i].byte_stream() == single.byte_stream()
        want = orc.jpeg_decompression(orc.jpeg_compression(rgb[i]))
        assert np.array_equal(out[i], want)
    bc.encode_device()
    bc.decode_device()
    dev = bc.d_out.download(np.uint8, out.size).reshape(out.shape)
    assert np.array_equal(dev, out)
    bc.close()


def test_decode_from_pageable_and_page_locked_tables():
    """The small transfers of the decoder go through an SM copy kernel: directly from page-locked host
    arrays (the batch codec's staging), through the plan's staging area from pageable ones.  Both decode
    to the same pixels, repeatedly (the staging area is reset after every synchronisation)."""
    from hiccup_b200 import entropy
    from hiccup_b200.batch import DctBatchCodec
    n, h, w = 3, 88, 120
    rgb = np.stack([orc.synthetic_image(h, w, 90 + i) for i in range(n)])
    bc = DctBatchCodec(n, h, w)
    enc = bc.encode(rgb)                               # tables and payloads are views of page-locked staging
    want = bc.decode(enc).copy()
    pageable = entropy.EncodedStreams(enc.layout, enc.index.copy(), enc.nsym.copy(), enc.nbits.copy(), enc.byte_off.copy(),
                                      enc.byte_len.copy(), enc.symbols.copy(), enc.packed.copy(), enc.data.copy())
    for _ in range(3):
        assert np.array_equal(bc.decode(pageable), want)
        assert np.array_equal(bc.decode(enc), want)
    for i in range(n):
        assert np.array_equal(want[i], orc.jpeg_decompression(orc.jpeg_compression(rgb[i])))
    bc.close()


def test_encoder_twice_on_one_plan_resets_its_histograms():
    """The compaction resets the histogram bins it reads, so a plan encodes batch after batch without a
    memset over the histogram arrays; a second, different batch must not see counts of the first."""
    from hiccup_b200.batch import DctBatchCodec
    n, h, w = 2, 72, 104
    a = np.stack([orc.synthetic_image(h, w, 400 + i) for i in range(n)])
    b = np.stack([orc.synthetic_image(h, w, 500 + i) for i in range(n)])
    bc = DctBatchCodec(n, h, w)
    fresh = DctBatchCodec(n, h, w)
    bc.encode(a)
    got = [im.byte_stream() for im in bc.hic_images(bc.encode(b))]
    want = [im.byte_stream() for im in fresh.hic_images(fresh.encode(b))]
    assert got == want
    for host_codes in (True, False, True):             # the host builder and the device builder interleaved
        bc.device_codes = not host_codes
        assert [im.byte_stream() for im in bc.hic_images(bc.encode(a if host_codes else b))] == \
               [im.byte_stream() for im in fresh.hic_images(fresh.encode(a if host_codes else b))]
    bc.close()
    fresh.close()


@pytest.mark.parametrize("mode,shape", [("dct", (72, 104)), ("wavelet", (64, 96))])
def test_pipelined_codec_equals_unchunked(mode, shape):
    """Chunks over concurrent streams/threads produce the same bytes and pixels as one launch per batch."""
    from hiccup_b200.batch import DctBatchCodec, PipelinedCodec, WaveletBatchCodec
    n, (h, w) = 12, shape
    rgb = np.stack([orc.synthetic_image(h, w, 300 + i) for i in range(n)])
    whole = (DctBatchCodec if mode == "dct" else WaveletBatchCodec)(n, h, w)
    enc = whole.encode(rgb)
    want_hic = [im.byte_stream() for im in whole.hic_images(enc)]
    want = whole.decode(enc).copy()
    pipe = PipelinedCodec(n, h, w, chunk=2, slots=3, mode=mode)
    got_hic = [None] * n

    def on_encoded(first, e):
        codec = pipe.codecs[0]
        for k, im in enumerate(codec.hic_images(e)):
            got_hic[first + k] = im.byte_stream()

    out = np.empty(pipe.out_shape, np.uint8)
    for from_device in (True, False, True):          # staging buffers are reused; both decode inputs
        out[...] = 0
        pipe.round_trip(rgb, out, on_encoded=on_encoded, from_device=from_device)
        assert np.array_equal(out, want)
        assert got_hic == want_hic
    pipe.close()
    whole.close()


@pytest.mark.parametrize("threads", [None, 2, 1])
def test_resident_multi_stream_steps_equal_the_single_codec(threads):
    """PipelinedCodec.device_steps (the batch parked on the device as one chunk per slot, every slot on its own CUDA
    stream, driven by `threads` host threads) leaves in every slot what the single codec leaves for those images."""
    from hiccup_b200.batch import DctBatchCodec, PipelinedCodec
    n, h, w, slots = 12, 72, 104, 4
    rgb = np.stack([orc.synthetic_image(h, w, 500 + i) for i in range(n)])
    whole = DctBatchCodec(n, h, w)
    enc = whole.encode(rgb)
    want = whole.decode(enc).copy()
    pipe = PipelinedCodec(n, h, w, chunk=n // slots, slots=slots)
    pipe.upload_resident(rgb)
    pipe.device_steps(2, threads=threads)
    per = n // slots
    for s_, codec in enumerate(pipe.codecs):
        got = codec.d_out.download(np.uint8, per * want[0].size).reshape((per,) + want.shape[1:])
        assert np.array_equal(got, want[s_ * per:(s_ + 1) * per]), "slot %d" % s_
        coef = codec.coefficients()
        assert np.array_equal(coef, whole.coefficients()[s_ * per:(s_ + 1) * per])
    pipe.close()
    whole.close()


def test_file_driver_writes_reference_files(tmp_path):
    """run.compress / run.decompress (reference run.py:18-43): the .hic file is byte-identical to the one
    the reference pickles, single file and directory (batched) mode alike."""
    import cv2
    from hiccup_b200 import run, model
    names = ["syn64", "syn48x80", "flat32"]
    paths = []
    for nm in names:
        g = load_golden(nm)
        p = str(tmp_path / (nm + ".png"))
        cv2.imwrite(p, g["rgb"])                      # imread returns exactly these bytes (PNG is lossless)
        paths.append(p)
    for nm, p in zip(names, paths):
        g = load_golden(nm)
        out = run.compress(p, str(tmp_path), model.Compression.JPEG)
        want = pickle.dumps(pickle.loads(g["hic"].tobytes()))
        assert open(out, "rb").read() == want
        if str(g["decode_error"]) == "":
            assert np.array_equal(run.decompress(out), g["rgb_out"])
    many = tmp_path / "many"
    many.mkdir()
    outs = run.compress_many(paths + [paths[0]], str(many), model.Compression.JPEG)
    assert len(outs) == 4
    for nm, p in zip(names, paths):
        g = load_golden(nm)
        got = open(str(many / (nm + ".png.JPEG-hic")), "rb").read()
        assert got == pickle.dumps(pickle.loads(g["hic"].tobytes()))
    w = load_golden("w_syn64")
    pw = str(tmp_path / "w.png")
    cv2.imwrite(pw, w["rgb"])
    out = run.compress(pw, str(tmp_path), model.Compression.HIC)
    assert open(out, "rb").read() == pickle.dumps(pickle.loads(w["hic"].tobytes()))
    assert np.array_equal(run.decompress(out), w["rgb_out"])


@pytest.mark.parametrize("mode", ["dct", "wavelet"])
def test_batch_files_round_trip(mode, tmp_path):
    """batch.hic_files / streams_from_files (the library's host threads) on a real encode: the files are the ones the
    per-image container path pickles, and reading them back decodes to the same pixels."""
    from hiccup_b200 import hicimage, model, run
    from hiccup_b200.batch import DctBatchCodec, WaveletBatchCodec
    h, w, n = (136, 200, 6) if mode == "dct" else (64, 96, 4)
    rgb = np.stack([orc.synthetic_image(h, w, 300 + i) for i in range(n)])
    rgb[-1] = 77                                                    # a flat image: one-symbol tables, shortest bit strings
    codec = (DctBatchCodec if mode == "dct" else WaveletBatchCodec)(n, h, w)
    enc = codec.encode(rgb)
    want_pixels = codec.decode(enc).copy()
    want = [pickle.dumps(hi.byte_stream()) for hi in codec.hic_images(enc)]
    assert hicimage._native().files_ok
    files = codec.hic_files(enc)
    assert [bytes(f) for f in files] == want
    back = codec.streams_from_files(files)
    assert np.array_equal(codec.decode(back), want_pixels)
    paths = [str(tmp_path / ("%d.hic" % i)) for i in range(n)]
    codec.write_files(enc, paths)
    assert np.array_equal(codec.decode(codec.read_files(paths)), want_pixels)
    codec.close()
    # the file driver: same pixels as one file at a time, in the order asked for
    order = [3, 0, 2]
    many = run.decompress_many([paths[i] for i in order])
    for i, got in zip(order, many):
        assert np.array_equal(got, want_pixels[i]) and np.array_equal(got, run.decompress(paths[i]))
    style = model.Compression.JPEG if mode == "dct" else model.Compression.HIC
    assert run._peek(want[0]) == (style, (h, w))


def test_argument_checks_of_the_new_entry_points():
    """hic_dct_tie_capacity sizes the tie buffer; a smaller one is refused (HIC_ERR_CAPACITY), and a bit stream that
    would run past the bytes the caller declared is refused before any kernel reads it."""
    import ctypes
    from hiccup_b200 import _lib, codec, compression, entropy
    lib = _lib.load()
    h, w = 64, 96
    g = _lib.geometry(h, w)
    cap = _lib.tie_capacity(1, h, w)
    assert cap > g.blocks_per_image
    rgb = orc.synthetic_image(h, w, 77)
    d_rgb = _lib.DeviceBuffer(rgb.nbytes)
    d_rgb.upload(rgb)
    coef = _lib.DeviceBuffer(g.blocks_per_image * 128)
    ties = _lib.DeviceBuffer(cap * _lib.TIE_RECORD_BYTES)
    stats = _lib.DeviceBuffer(4 * _lib.TIE_STATS)
    assert lib.hic_dct_forward(d_rgb.ptr, 1, h, w, coef.ptr, ties.ptr, 1, stats.ptr, None) == -3          # HIC_ERR_CAPACITY
    assert lib.hic_dct_forward(d_rgb.ptr, 1, h, w, coef.ptr, ties.ptr, cap, stats.ptr, None) == 0
    _lib.sync()
    hi = codec.jpeg_encode(compression.jpeg_compression(rgb))
    p = hi.payloads
    order = [kind * 3 + c for c in range(3) for kind in range(3)]
    rows, syms, lens, codes = codec._tables_to_arrays([p[i] for i in order])
    data, offs, nbits = codec._gather_payload_bytes([p[9 + i] for i in order])
    dec = entropy.EntropyDecoder(_lib.layout_dct(1, h, w))
    try:
        dec.decode(rows, syms, lens, codes, data, offs, nbits, coef.ptr)                                 # fits
        with pytest.raises(_lib.HicError):
            dec.decode(rows, syms, lens, codes, data[:len(data) // 2], offs, nbits, coef.ptr)           # declared too short
    finally:
        dec.close()
        for b in (d_rgb, coef, ties, stats):
            b.free()
