"""hiccup_b200 -- B200-native encode/decode hot path for the hiccup image codec.

Drop-in mirror of the reference's entry points (reference file:line in each module):

    hiccup_b200.compression   jpeg_compression / jpeg_decompression / wavelet_compression / wavelet_decompression
    hiccup_b200.codec         jpeg_encode / jpeg_decode / wavelet_encode / wavelet_decode
    hiccup_b200.hicimage      HicImage and payload classes (the `.hic` wire format)
    hiccup_b200.batch         additive batched / device-resident entry points

All numeric work runs in hand-written sm_100a CUDA kernels reached through the C ABI declared in
include/hiccup_b200.h (hiccup_b200/csrc -> libhiccup_b200.so, loaded with ctypes).  There is no
CPU fallback: importing the numeric modules without the built library raises.
"""
from hiccup_b200 import model, hicimage, iohelper, settings, _compat  # noqa: F401

_compat.resolve()

__all__ = ["model", "hicimage", "iohelper", "settings"]
__version__ = "0.1.0"
