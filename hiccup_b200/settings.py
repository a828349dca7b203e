"""Module-level configuration, same names and defaults as reference hiccup/settings.py:10-23.

Read at call time by hiccup_b200.compression / hiccup_b200.codec.  The CUDA path implements the
defaults below; a setting the kernels do not implement raises NotImplementedError instead of
silently computing something else.
"""
from hiccup_b200 import model

DEBUG = False

WAVELET = model.Wavelet.DAUBECHIE
WAVELET_QUALITY_FACTOR = 1
WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER = 1
WAVELET_THRESHOLD = 5
WAVELET_NUM_LEVELS = 3
WAVELET_TILES = 8          # dead in the reference too (settings.py:17 has no readers)

JPEG_BLOCK_SIZE = 8


def JPEG_BLOCK_SHAPE():
    return JPEG_BLOCK_SIZE, JPEG_BLOCK_SIZE


def check_supported():
    if JPEG_BLOCK_SIZE != 8:
        raise NotImplementedError("the CUDA DCT path is built for 8x8 blocks (settings.JPEG_BLOCK_SIZE=%r)"
                                  % (JPEG_BLOCK_SIZE,))


def check_wavelet_supported():
    if WAVELET not in (model.Wavelet.DAUBECHIE, model.Wavelet.HAAR):
        raise NotImplementedError("the CUDA wavelet path implements db1/haar only (settings.WAVELET=%r)" % (WAVELET,))
    if WAVELET_NUM_LEVELS != 3:
        raise NotImplementedError("the CUDA wavelet path implements 3 levels (settings.WAVELET_NUM_LEVELS=%r)"
                                  % (WAVELET_NUM_LEVELS,))
    if WAVELET_QUALITY_FACTOR != 1:
        raise NotImplementedError("WAVELET_QUALITY_FACTOR != 1 (an order statistic over the channel) is not on the CUDA path")
    if WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER != 1 or WAVELET_THRESHOLD != 5:
        raise NotImplementedError("the CUDA wavelet path implements multiplier 1 and threshold 5")
