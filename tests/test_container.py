"""The `.hic` container and bit framing (host only): byte-for-byte round trips of reference files."""
import glob
import os
import pickle

import numpy as np
import pytest

from tests.conftest import GOLDEN, load_golden

ALL = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "*.npz")))


def test_iohelper_kats():
    from hiccup_b200 import iohelper
    assert iohelper.padded_bs_2_bytes("101") == b"\x05\xa0"           # iohelpertest.py:11-13
    assert iohelper.padded_bytes_2_bs(bytearray(b"\x05\xa0")) == "101"  # iohelpertest.py:15-17
    assert iohelper.padded_bs_2_bytes("10010110") == b"\x08\x96\x00"  # aligned input gets a whole zero byte
    for s in ["01", "0000", "1010000", "00000", "000111", "00000001", "000001", "0000001", "10010110", "0" * 901 + "1"]:
        assert iohelper.padded_bytes_2_bs(iohelper.padded_bs_2_bytes(s)) == s
    rng = np.random.default_rng(0)
    for _ in range(100):
        s = bin(int(rng.integers(0, 100000)))[2:]
        assert iohelper.padded_bytes_2_bs(iohelper.padded_bs_2_bytes(s)) == s


@pytest.mark.parametrize("name", ALL)
def test_reference_files_round_trip_byte_for_byte(name):
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import hicimage, model
    g = load_golden(name)
    stream = pickle.loads(g["hic"].tobytes())
    hi = hicimage.HicImage.from_bytes(stream)
    want_type = model.Compression.JPEG if str(g["mode"]) == "jpeg" else model.Compression.HIC
    assert hi.hic_type == want_type
    assert len(hi.payloads) == (20 if want_type == model.Compression.JPEG else 14)
    assert hi.byte_stream() == stream
    # rebuilding every payload from its logical content gives the same bytes (hicimagetest.py:9-36)
    n_tab = 9 if want_type == model.Compression.JPEG else 6
    rebuilt = ([hicimage.PayloadStringP.from_rows(p.rows) for p in hi.payloads[:n_tab]]
               + [hicimage.BitStringP(p.payload) for p in hi.payloads[n_tab:2 * n_tab]]
               + [hicimage.TupP(*p.numbers) for p in hi.payloads[2 * n_tab:]])
    again = hicimage.HicImage(hi.hic_type, hi.settings, rebuilt)
    assert again.byte_stream() == stream


def test_write_and_read_file(tmp_path):
    import hiccup_b200  # noqa: F401
    from hiccup_b200 import hicimage
    g = load_golden("syn64")
    stream = pickle.loads(g["hic"].tobytes())
    hi = hicimage.HicImage.from_bytes(stream)
    path = os.path.join(tmp_path, "a.hic")
    hi.write_file(path)
    assert hicimage.HicImage.from_file(path).byte_stream() == stream
    with open(path, "rb") as f:
        assert pickle.load(f) == stream          # the file is exactly what the reference writes (hicimage.py:171-174)


def test_payload_equality_semantics():
    from hiccup_b200 import hicimage
    assert hicimage.TupP(1, "1") == hicimage.TupP(1, "1")
    assert hicimage.BitStringP("101") == hicimage.BitStringP.from_bytes(b"\x05\xa0")
    assert hicimage.BitStringP("101").bit_count == 3 == hicimage.BitStringP.from_bytes(b"\x05\xa0").bit_count
