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
    """Raise for the wavelet settings the CUDA path does not implement: the coif1 / sym2 filters (their
    taps come from PyWavelets, which is not available to restate or test against), a multiplier of 0
    (the reference then thresholds unrounded floats through np.vectorize, whose output type depends on the
    first element), and a quality factor outside (0, 1]."""
    if WAVELET not in (model.Wavelet.DAUBECHIE, model.Wavelet.HAAR):
        raise NotImplementedError("the CUDA wavelet path implements db1/haar only (settings.WAVELET=%r)" % (WAVELET,))
    if int(WAVELET_NUM_LEVELS) != WAVELET_NUM_LEVELS or not 1 <= WAVELET_NUM_LEVELS <= 5:
        raise NotImplementedError("the CUDA wavelet path implements 1..5 levels (settings.WAVELET_NUM_LEVELS=%r)"
                                  % (WAVELET_NUM_LEVELS,))
    if not 0 < WAVELET_QUALITY_FACTOR <= 1:
        raise ValueError("WAVELET_QUALITY_FACTOR must be in (0, 1] (got %r)" % (WAVELET_QUALITY_FACTOR,))
    if WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER == 0:
        raise NotImplementedError("WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER = 0 (no quantisation) is not on the CUDA path")
    if WAVELET_THRESHOLD < 0:
        raise ValueError("WAVELET_THRESHOLD must be >= 0 (got %r)" % (WAVELET_THRESHOLD,))


def wavelet_defaults():
    """True when the fused default-settings kernels (K9 / K10) apply; otherwise the general path runs."""
    return (WAVELET_NUM_LEVELS == 3 and WAVELET_QUALITY_FACTOR == 1 and WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER == 1
            and WAVELET_THRESHOLD == 5)
