"""The oracle against every known-answer vector the reference's own tests hold for the hot path
(reference hiccup/test/*.py)."""
import numpy as np

from oracle import hiccup_oracle as orc


def test_zigzag_order_kat():
    # transformtest.py:96-105: 3x3 [[1,2,3],[4,5,6],[7,8,9]] -> 1,4,2,3,5,7,8,6,9
    m = np.arange(1, 10).reshape(3, 3)
    assert list(orc.zigzag(m)) == [1, 4, 2, 3, 5, 7, 8, 6, 9]
    assert np.array_equal(orc.izigzag(orc.zigzag(m), (3, 3)), m)


def test_zigzag8_table_matches_survey():
    want = [0, 8, 1, 2, 9, 16, 24, 17, 10, 3, 4, 11, 18, 25, 32, 40, 33, 26, 19, 12, 5, 6, 13, 20, 27, 34, 41, 48,
            56, 49, 42, 35, 28, 21, 14, 7, 15, 22, 29, 36, 43, 50, 57, 58, 51, 44, 37, 30, 23, 31, 38, 45, 52,
            59, 60, 53, 46, 39, 47, 54, 61, 62, 55, 63]
    assert list(orc.ZIGZAG8) == want


def test_dc_difference_kat():
    # codectest.py:14-18: 8x8 of i+k split into 4x4 blocks -> [0, 4, 0, 4]
    m = np.array([[i + k for k in range(8)] for i in range(8)])
    dc = orc.split_blocks(m, 4)[:, 0, 0]
    assert list(orc.differences(dc)) == [0, 4, 0, 4]


def test_run_length_kats():
    # codectest.py:20-35
    m = np.zeros((8, 8), dtype=np.int64)
    m[0, :4] = [99, -59, 0, 7]
    m[4, :2] = [12, -2]
    lengths, values = orc.run_length(orc.zigzag(m))
    assert list(zip(lengths, values)) == [(0, 99), (1, -59), (6, 7), (4, 12), (1, -2), (0, 0)]
    # codectest.py:37-47
    lengths, values = orc.run_length(orc.zigzag(np.array([[1, 2], [3, 4]]))[1:])
    assert list(zip(lengths, values)) == [(0, 3), (0, 2), (0, 4)]
    # codectest.py:49-54: 17 zeros then 1 -> lengths [14, 2]
    lengths, values = orc.run_length(np.array([0] * 17 + [1]))
    assert list(lengths) == [14, 2] and list(values) == [0, 1]
    # codectest.py:56-67: trailing (0, 0) rule
    l1, v1 = orc.run_length(np.array([0, 0, 5]))
    assert not (l1[-1] == 0 and v1[-1] == 0)
    l2, v2 = orc.run_length(np.array([0, 0, 5, 0, 0]))
    assert l2[-1] == 0 and v2[-1] == 0


def test_run_length_inverse_random():
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(1, 400))
        a = (rng.integers(-5, 6, n) * (rng.random(n) < 0.2)).astype(np.int64)
        lengths, values = orc.run_length(a)
        assert np.array_equal(orc.decode_run_length(lengths, values, n), a)


def test_huffman_kats():
    # huffmantest.py:11-15 singleton; :51-58 1->001 2->000 3->01 4->1
    assert orc.huffman_table([0, 0, 0, 0, 0]) == [(0, "1")]
    t = orc.huffman_table([1, 2, 2, 3, 3, 3, 4, 4, 4, 4])
    assert [(int(s), c) for s, c in t] == [(1, "001"), (2, "000"), (3, "01"), (4, "1")]
    bits = orc.huffman_encode_bits([1, 2, 2, 3, 3, 3, 4, 4, 4, 4], t)
    assert bits == "001" + "000" * 2 + "01" * 3 + "1" * 4
    assert [int(v) for v in orc.huffman_decode_bits(bits, t)] == [1, 2, 2, 3, 3, 3, 4, 4, 4, 4]
    # huffmantest.py:17-32: smaller frequency on the left ('1')
    t2 = dict((int(s), c) for s, c in orc.huffman_table([0, 0, 0, 1]))
    assert t2 == {0: "0", 1: "1"}


def test_bit_padding_kat():
    # iohelpertest.py:11-17
    assert orc.padded_bits_to_bytes("101") == b"\x05\xa0"
    assert orc.padded_bytes_to_bits(b"\x05\xa0") == "101"
    for s in ["01", "0000", "1010000", "00000", "000111", "00000001", "000001", "0000001", "10010110", "0" * 901 + "1"]:
        assert orc.padded_bytes_to_bits(orc.padded_bits_to_bytes(s)) == s


def test_jpeg_encode_layout_kat():
    # codectest.py:69-79: 2x2 planes -> 20 payloads, first table's first row is (1, '1')
    planes = {"lum": np.array([[1, 2], [3, 4]]), "cr": np.array([[5, 6], [7, 8]]), "cb": np.array([[9, 10], [11, 12]])}
    enc = orc.jpeg_encode(planes)
    assert len(enc["tables"]) + len(enc["bits"]) + len(enc["shapes"]) == 20
    assert (int(enc["tables"][0][0][0]), enc["tables"][0][0][1]) == (1, "1")
    dec = orc.jpeg_decode(enc)
    for ch in orc.CHANNELS:
        assert np.array_equal(dec[ch], planes[ch])


def test_quantisation_kats():
    # quantizationtest.py:12-25: a block of 16s divided by the luminance table -> 1 at (0,0)
    q = orc.np_round_i32(np.divide(np.full((8, 8), 16.0), orc.LUM_TABLE))
    assert q[0, 0] == 1
    # quantizationtest.py:32-43: np.round is half-to-even
    assert list(orc.np_round_i32(np.array([0.5, 1.5, 2.5, -0.5]))) == [0, 2, 2, 0]


def test_dct_channel_of_128_is_zero():
    # transformtest.py:148-151
    assert not orc.dct_channel(np.full((16, 16), 128, np.uint8), orc.LUM_TABLE).any()


def test_padding_and_blocks():
    # transformtest.py:10-94
    m = np.arange(30).reshape(5, 6)
    p = orc.pad_matrix(m, 4)
    assert p.shape == (8, 8) and np.array_equal(p[:5, :6], m) and not p[5:].any() and not p[:, 6:].any()
    b = orc.split_blocks(m, 4)
    assert b.shape == (4, 4, 4)
    assert np.array_equal(orc.merge_blocks(b, (5, 6), 4), m)
