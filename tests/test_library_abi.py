"""The C-ABI library loads and exports every symbol include/hiccup_b200.h declares (no compute
calls: this runs without a GPU), and the product fails loudly when there is no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "hiccup_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hic_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hiccup_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name


def test_binding_table_covers_the_header():
    from hiccup_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_host_only_entry_points_work_without_a_device():
    from hiccup_b200 import _lib
    lib = _lib.load()
    assert lib.hic_version() >= 100
    g = _lib.geometry(426, 640)
    assert (g.hc, g.wc, g.nby_l, g.nbx_l, g.nby_c, g.nbx_c) == (213, 320, 54, 80, 27, 40)
    assert (g.nb_l, g.nb_c, g.blocks_per_image, g.out_h, g.out_w) == (4320, 1080, 6480, 426, 640)
    lay = _lib.layout_dct(2, 26, 40)
    assert list(lay.nb) == [20, 6, 6] and list(lay.len) == [1260, 378, 378] and lay.skip_first == 1
    with pytest.raises(_lib.HicError):
        _lib.geometry(1, 1)


def test_no_cpu_fallback():
    """Without a CUDA device the numeric entry points raise; they never compute on the host."""
    import numpy as np
    from hiccup_b200 import _lib, compression
    n = ctypes.c_int(0)
    if _lib.load().hic_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.HicError):
        compression.jpeg_compression(np.zeros((16, 16, 3), np.uint8))


def test_product_and_tools_never_import_the_oracle():
    """oracle/ is test infrastructure: nothing under hiccup_b200/ (the product) or tools/ may import it."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)", re.M)
    for path in glob.glob(os.path.join(root, "hiccup_b200", "**", "*.py"), recursive=True) + glob.glob(os.path.join(root, "tools", "*.py")):
        with open(path) as f:
            assert not pat.search(f.read()), path
