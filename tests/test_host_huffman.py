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
