"""The batch `.hic` file writer (hic_hicfile_pack_files, host threads of the library; batch.hic_files) against the Python
container path, which is itself pinned to the reference's bytes (tests/test_container.py): same files, byte for byte.
No GPU: the encode results come from the oracle."""
import pickle

import numpy as np
import pytest

from oracle import hiccup_oracle as orc


def _encoded_streams(per_image, mode):
    """EncodedStreams (the batch codecs' host-side result) from the oracle's tables and bit strings of each image."""
    from hiccup_b200 import _lib, entropy
    n = len(per_image)
    rows, symbols, lens, codes, chunks, byte_off, byte_len, nbits = [], [], [], [], [], [], [], []
    pos = 0
    for enc in per_image:
        for c in range(3):
            for kind in range(3):
                if mode == "dct":
                    k = kind * 3 + c
                elif kind == entropy.KIND_DC:
                    k = None
                else:
                    k = (kind - 1) * 3 + c
                table = [] if k is None else enc["tables"][k]
                framed = b"" if k is None else orc.padded_bits_to_bytes(enc["bits"][k])
                rows.append(len(table))
                symbols += [int(a) for a, _ in table]
                lens += [len(b) for _, b in table]
                codes += [int(b, 2) for _, b in table]
                byte_off.append(pos)
                byte_len.append(len(framed))
                nbits.append(0 if k is None else len(enc["bits"][k]))
                chunks.append(framed + b"\0" * (-len(framed) % 4))
                pos += len(chunks[-1])
    data = np.frombuffer(b"".join(chunks) or b"\0", np.uint8)
    layout = _lib.StreamLayout()
    layout.n_images = n
    return entropy.EncodedStreams.from_tables(layout, rows, None, np.array(nbits, np.uint64), np.array(byte_off, np.uint64),
                                              np.array(byte_len, np.uint64), np.array(symbols, np.int32), np.array(lens, np.uint8),
                                              np.array(codes, np.uint64), data)


def _stub(cls, n, g):
    codec = object.__new__(cls)
    codec.n, codec.g = n, g
    return codec


@pytest.mark.parametrize("threads", [1, 3])
def test_dct_batch_files_equal_the_python_container(threads):
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import _lib, hicimage
    from hiccup_b200.batch import DctBatchCodec
    assert hicimage._native().files_ok, "the native file writer did not calibrate here (it must fall back, not fail)"
    h, w = 136, 200
    images = [orc.synthetic_image(h, w, 70 + i) for i in range(4)]
    images.append(np.full((h, w, 3), 77, np.uint8))                  # flat: every AC stream is the end marker alone
    images.append(np.random.default_rng(5).integers(0, 256, (h, w, 3), dtype=np.uint8))     # noise: the biggest tables
    per_image = [orc.jpeg_encode(orc.jpeg_compression(im)) for im in images]
    enc = _encoded_streams(per_image, "dct")
    codec = _stub(DctBatchCodec, len(images), _lib.geometry(h, w))
    want = [pickle.dumps(hi.byte_stream()) for hi in codec.hic_images(enc)]
    got = codec.hic_files(enc, threads=threads)
    assert [bytes(b) for b in got] == want
    # the oracle's own container bytes, and a subset in another order
    for i, e in enumerate(per_image):
        stream = pickle.loads(bytes(got[i]))
        assert stream[0] == b"JPEG" and len(stream) == 21
        assert stream[10:19] == [orc.padded_bits_to_bytes(b) for b in e["bits"]]
        back = hicimage.HicImage.from_bytes(stream)
        assert [[(int(a), c) for a, c in p.rows] for p in back.payloads[:9]] == [[(int(a), c) for a, c in t] for t in e["tables"]]
    sub = codec.hic_files(enc, images=[5, 0, 2], threads=threads)
    assert [bytes(b) for b in sub] == [want[5], want[0], want[2]]
    assert codec.hic_files(enc, images=[]) == []
    # reuse=True: the same bytes, written into a buffer the codec keeps between calls
    first = codec.hic_files(enc, threads=threads, reuse=True)
    assert [bytes(b) for b in first] == want
    keep = codec._files_keep["out"]
    again = codec.hic_files(enc, images=[1, 4], threads=threads, reuse=True)
    assert codec._files_keep["out"] is keep and [bytes(b) for b in again] == [want[1], want[4]]


def test_dct_batch_files_big_image(tmp_path):
    """640x426: tables past the 64 KiB frame size and bit strings that pickle writes outside its frames; through
    write_files and back through HicImage.from_file."""
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import _lib, hicimage
    from hiccup_b200.batch import DctBatchCodec
    h, w = 426, 640
    rgb = orc.synthetic_image(h, w, 5)
    e = orc.jpeg_encode(orc.jpeg_compression(rgb))
    enc = _encoded_streams([e, e], "dct")
    codec = _stub(DctBatchCodec, 2, _lib.geometry(h, w))
    want = pickle.dumps(codec.hic_images(enc, images=[1])[0].byte_stream())
    assert max(len(b) for b in pickle.loads(want)) > 65536
    paths = [str(tmp_path / "a.hic"), str(tmp_path / "b.hic")]
    codec.write_files(enc, paths)
    for p in paths:
        with open(p, "rb") as f:
            assert f.read() == want
        hi = hicimage.HicImage.from_file(p)
        assert hi.byte_stream() == pickle.loads(want)


def test_wavelet_batch_files_equal_the_python_container():
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import _lib
    from hiccup_b200.batch import WaveletBatchCodec
    h, w = 64, 96
    images = [orc.synthetic_image(h, w, 90 + i) for i in range(3)] + [np.full((h, w, 3), 10, np.uint8)]
    per_image = [orc.wavelet_encode(orc.wavelet_compression(im)) for im in images]
    enc = _encoded_streams(per_image, "wavelet")
    codec = _stub(WaveletBatchCodec, len(images), _lib.wavelet_geometry(h, w))
    want = [pickle.dumps(hi.byte_stream()) for hi in codec.hic_images(enc)]
    got = codec.hic_files(enc, threads=2)
    assert [bytes(b) for b in got] == want
    assert pickle.loads(want[0])[0] == b"HIC" and len(pickle.loads(want[0])) == 15


def test_batch_files_fall_back_when_the_library_declines(monkeypatch):
    """Entries shorter than two bytes are left to the Python path by the library (CPython shares such objects, so pickle
    may write memo references); a writer that did not calibrate leaves everything to it.  Same bytes either way."""
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import _lib, hicimage
    from hiccup_b200.batch import DctBatchCodec
    h, w = 48, 80
    e = orc.jpeg_encode(orc.jpeg_compression(orc.synthetic_image(h, w, 3)))
    enc = _encoded_streams([e, e, e], "dct")
    codec = _stub(DctBatchCodec, 3, _lib.geometry(h, w))
    want = [pickle.dumps(hi.byte_stream()) for hi in codec.hic_images(enc)]
    enc.byte_len = enc.byte_len.copy()
    keep = int(enc.byte_len[9 + 4])
    enc.byte_len[9 + 4] = 1                                          # image 1, one bit string cut to a single byte
    nat = hicimage._native()
    stream_of, modes, lead, trail = codec._file_plan([0, 1, 2])
    res = nat.pack_files(hiccup_b200._compat.wire_tuple_class(), enc.index, enc.symbols, enc.packed, enc.data, enc.byte_off, enc.byte_len,
                         stream_of, modes, lead, trail, 2)
    assert [int(v) > 0 for v in res[2]] == [True, False, True]
    got = codec.hic_files(enc)
    assert bytes(got[0]) == want[0] and bytes(got[2]) == want[2]
    assert bytes(got[1]) == pickle.dumps(codec.hic_images(enc, images=[1])[0].byte_stream())
    enc.byte_len[9 + 4] = keep
    monkeypatch.setattr(nat, "files_ok", False)
    assert [bytes(b) for b in codec.hic_files(enc)] == want


def _same_streams(a, b, n_streams):
    assert np.array_equal(a.index[:, 1], b.index[:, 1])
    for s in range(n_streams):
        sa, sb = a.stream_rows(s), b.stream_rows(s)
        assert all(np.array_equal(x, y) for x, y in zip(sa, sb)), s
        assert a.framed(s) == b.framed(s), s
        assert int(a.nbits[s]) == int(b.nbits[s]), s
        assert int(a.byte_off[s]) % 4 == 0


@pytest.mark.parametrize("mode", ["dct", "wavelet"])
def test_files_read_back_to_the_same_streams(mode, tmp_path):
    """streams_from_files (hic_hicfile_scan_files / _parse_files) inverts hic_files; files in another pickle protocol take
    the unpickler and give the same streams; files of another shape or mode are refused."""
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import _lib, hicimage
    from hiccup_b200.batch import DctBatchCodec, WaveletBatchCodec
    h, w = 72, 104
    images = [orc.synthetic_image(h, w, 30 + i) for i in range(5)]
    if mode == "dct":
        per_image = [orc.jpeg_encode(orc.jpeg_compression(im)) for im in images]
        codec = _stub(DctBatchCodec, 5, _lib.geometry(h, w))
        other = _stub(DctBatchCodec, 5, _lib.geometry(h + 8, w))
    else:
        per_image = [orc.wavelet_encode(orc.wavelet_compression(im)) for im in images]
        codec = _stub(WaveletBatchCodec, 5, _lib.wavelet_geometry(h, w))
        other = _stub(WaveletBatchCodec, 5, _lib.wavelet_geometry(h + 8, w))
    codec.layout = other.layout = None
    enc = _encoded_streams(per_image, mode)
    files = codec.hic_files(enc, threads=2)
    nat = hicimage._native()
    calls = []
    real = nat.parse_files
    nat.parse_files = lambda *a, **k: calls.append(1) or real(*a, **k)
    try:
        back = codec.streams_from_files(files, threads=3)
        assert calls and real(files, codec._file_plan(list(range(5)))[0], 45, 2, 2) is not None      # the native path took them
        _same_streams(back, enc, 45)
        # other protocols: the tolerant path
        old = [pickle.dumps(pickle.loads(bytes(f)), protocol=2) for f in files]
        assert real(old, codec._file_plan(list(range(5)))[0], 45, 2, 2) is None
        _same_streams(codec.streams_from_files(old), enc, 45)
        # on disk
        paths = [str(tmp_path / ("%d.hic" % i)) for i in range(5)]
        codec.write_files(enc, paths)
        _same_streams(codec.read_files(paths), enc, 45)
        with pytest.raises(ValueError):
            other.streams_from_files(files)
        with pytest.raises(ValueError):
            other.streams_from_files(old)
        with pytest.raises(ValueError):
            codec.streams_from_files(files[:3])
        wrong = _stub(WaveletBatchCodec if mode == "dct" else DctBatchCodec, 5, codec.g)
        wrong.layout = None
        with pytest.raises((ValueError, AssertionError, IndexError, AttributeError)):
            wrong.streams_from_files(files)
    finally:
        nat.parse_files = real


def test_damaged_files_are_refused_or_read_identically():
    """Byte damage anywhere in a file: the native reader either declines (and the unpickler decides) or returns exactly
    what the unpickler path returns."""
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import _lib, hicimage
    from hiccup_b200.batch import DctBatchCodec
    h, w = 40, 56
    e = orc.jpeg_encode(orc.jpeg_compression(orc.synthetic_image(h, w, 8)))
    enc = _encoded_streams([e], "dct")
    codec = _stub(DctBatchCodec, 1, _lib.geometry(h, w))
    codec.layout = None
    good = bytes(codec.hic_files(enc)[0])
    nat = hicimage._native()
    plan = codec._file_plan([0])[0]
    rng = np.random.default_rng(2)
    accepted = declined = 0
    for trial in range(1500):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 3))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        b = bytes(b)
        res = nat.parse_files([b], plan, 9, 2, 1)
        if res is None:
            declined += 1
            continue
        accepted += 1
        try:
            slow = hicimage.HicImage.from_bytes(hicimage.loads(b))
        except Exception:
            continue                                        # the unpickler is stricter where no data lives (frame lengths)
        order = [kind * 3 + c for c in range(3) for kind in range(3)]
        for s, k in enumerate(order):
            want = slow.payloads[k].arrays()
            a, n = int(res[0][s, 0]), int(res[0][s, 1])
            assert want is not None and np.array_equal(res[1][a:a + n], want[0]), trial
            assert np.array_equal(res[2][a:a + n] >> np.uint64(58), want[1]) and np.array_equal(res[2][a:a + n] & np.uint64((1 << 58) - 1), want[2])
            assert res[3][int(res[4][s]):int(res[4][s] + res[5][s])].tobytes() == bytes(slow.payloads[9 + k].byte_stream)
    assert accepted > 0 and declined > 0
