"""The C++ heapq replay (csrc/hic_huffman.cuh) against the oracle's real heapq and the reference's KATs."""
import ctypes

import numpy as np

from oracle import hiccup_oracle as orc


def _build(freqs):
    from hiccup_b200 import _lib
    lib = _lib.load()
    f = np.ascontiguousarray(freqs, np.uint32)
    lens = np.zeros(len(f), np.uint8)
    codes = np.zeros(len(f), np.uint64)
    _lib.check(lib.hic_huffman_build_host(f.ctypes.data, len(f), lens.ctypes.data, codes.ctypes.data))
    return [format(int(c), "0%db" % int(l)) for l, c in zip(lens, codes)]


def test_reference_kats():
    # huffmantest.py:11-15 singleton; :51-58 code KAT 1->001, 2->000, 3->01, 4->1
    assert _build([5]) == ["1"]
    assert _build([1, 2, 3, 4]) == ["001", "000", "01", "1"]
    assert _build([3, 1]) == ["0", "1"]          # huffmantest.py:17-32: smaller frequency is left = '1'


def test_matches_real_heapq_with_many_ties():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [100, 257, 1000, 1431]:
        for hi in (2, 5, 1000):
            freqs = rng.integers(1, hi, n)
            assert _build(freqs) == orc.huffman_codes(freqs), (n, hi)


def test_geometric_frequencies_deep_tree():
    freqs = [2 ** i for i in range(31)]
    assert _build(freqs) == orc.huffman_codes(freqs)


# ---- the device builder's packed 32-bit replay (csrc/hic_replay.cuh), run on the host ---------------
def _harness():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu_harness", "libhic_cpu_harness.so")
    import __graft_entry__
    __graft_entry__.build()
    return ctypes.CDLL(path)


def _codes_from_parents(par, n):
    """What huffman_codes_kernel reads off the parent links: leaf -> root, first popped (left) = '1'."""
    root = 2 * n - 2
    out = []
    for i in range(n):
        node, bits = i, ""
        while node != root:
            p = int(par[node])
            bits = ("1" if p & 0x8000 else "0") + bits
            node = p & 0x7FFF
        out.append(bits)
    return out


def _replay_narrow(freqs, mode):
    lib = _harness()
    f = np.ascontiguousarray(freqs, np.uint32)
    par = np.zeros(2 * len(f), np.uint16)
    rc = lib.hx_replay_narrow(f.ctypes.data_as(ctypes.c_void_p), len(f), mode, par.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    return _codes_from_parents(par, len(f))


def test_packed_replay_matches_real_heapq():
    rng = np.random.default_rng(1)
    for mode in (1, 2):
        for n in list(range(2, 70)) + [100, 255, 256, 257, 1000, 1431, 2047, 2048, 4099, 8192]:
            for hi in (2, 3, 5, 31):
                freqs = rng.integers(1, hi, n)
                assert _replay_narrow(freqs, mode) == orc.huffman_codes(freqs), (mode, n, hi)
        # skewed: a few heavy symbols and a long tail of ones (what DC-difference histograms mode like)
        for n in (300, 1737, 3000):
            freqs = np.ones(n, np.int64)
            freqs[:40] = rng.integers(1, 4000, 40)
            rng.shuffle(freqs)
            assert _replay_narrow(freqs, mode) == orc.huffman_codes(freqs), (mode, n)
        freqs = [2 ** i for i in range(17)]                       # sum 2^17 - 1: deepest tree that still packs
        assert _replay_narrow(freqs, mode) == orc.huffman_codes(freqs)


def test_packed_replay_refuses_what_does_not_pack():
    lib = _harness()
    f = np.array([1 << 17, 1 << 17], np.uint32)                   # sum = 2^18
    par = np.zeros(4, np.uint16)
    assert lib.hx_replay_narrow(f.ctypes.data_as(ctypes.c_void_p), 2, 1, par.ctypes.data_as(ctypes.c_void_p)) == 2
