"""Transform-stage entry points, drop-in for reference hiccup/compression.py.

    jpeg_compression(rgb)            compression.py:16-39
    jpeg_decompression(compressed)   compression.py:42-56
    wavelet_compression(rgb)         compression.py:59-85
    wavelet_decompression(c)         compression.py:88-100

Same argument meaning, return types and error behaviour; the arithmetic runs in the CUDA kernels
behind include/hiccup_b200.h (K1 forward, K7+K8 inverse, K9/K10 wavelet).
"""
import numpy as np

from hiccup_b200 import _lib, model, settings

LAST_STATS = {}        #: tie statistics of the most recent jpeg_compression call
LAST_INVERSE_STATS = {}  #: same for the most recent jpeg_decompression call


def _as_rgb(rgb):
    a = np.asarray(rgb)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError("expected an H x W x 3 image, got shape %r" % (a.shape,))
    if a.dtype != np.uint8:
        raise TypeError("expected uint8 pixels, got %s" % a.dtype)
    return np.ascontiguousarray(a)


def dct_forward_device(d_rgb, n, h, w, stream=None):
    """Run K1 on a device-resident batch.  Returns (coef buffer, stats array[4])."""
    lib = _lib.load()
    g = _lib.geometry(h, w)
    blocks = n * g.blocks_per_image
    coef = _lib.scratch(blocks * 128)
    capacity = _lib.tie_capacity(n, h, w)
    ties = _lib.scratch(capacity * _lib.TIE_RECORD_BYTES)
    stats = _lib.scratch(4 * _lib.TIE_STATS)
    _lib.check(lib.hic_dct_forward(d_rgb, n, h, w, coef.ptr, ties.ptr, capacity, stats.ptr, stream))
    st = stats.download(np.uint32, _lib.TIE_STATS, stream)
    if st[3]:
        raise _lib.HicError("tie list overflow (%d records dropped)" % st[3])
    ties.free()
    return coef, st


def jpeg_compression(rgb_image: np.ndarray) -> model.CompressedImage:
    """RGB -> YCrCb, chroma pyrDown, 8x8 DCT, Annex-K quantisation; int32 coefficient planes."""
    settings.check_supported()
    _lib.require_device()
    lib = _lib.load()
    img = _as_rgb(rgb_image)
    h, w = img.shape[:2]
    g = _lib.geometry(h, w)
    d_rgb = _lib.scratch(img.nbytes)
    d_rgb.upload(img)
    coef, st = dct_forward_device(d_rgb.ptr, 1, h, w)
    LAST_STATS.update(flagged_blocks=int(st[0]), reevaluated=int(st[1]), changed=int(st[2]))
    lum = _lib.scratch(4 * h * w)
    cr = _lib.scratch(4 * g.hc * g.wc)
    cb = _lib.scratch(4 * g.hc * g.wc)
    _lib.check(lib.hic_blocks_to_planes(coef.ptr, 1, h, w, lum.ptr, cr.ptr, cb.ptr, None))
    out = model.CompressedImage(lum.download(np.int32, h * w).reshape(h, w),
                                cr.download(np.int32, g.hc * g.wc).reshape(g.hc, g.wc),
                                cb.download(np.int32, g.hc * g.wc).reshape(g.hc, g.wc))
    for b in (d_rgb, coef, lum, cr, cb):
        b.free()
    return out


def _plane_i32(a):
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("coefficient plane must be 2-D, got shape %r" % (a.shape,))
    return np.ascontiguousarray(a.astype(np.int32))


def planes_to_device_blocks(compressed, stream=None):
    """Upload a CompressedImage's three planes and convert them to zigzag blocks on the device."""
    lib = _lib.load()
    d = compressed.as_dict
    lum, cr, cb = _plane_i32(d["lum"]), _plane_i32(d["cr"]), _plane_i32(d["cb"])
    h, w = lum.shape
    g = _lib.geometry(h, w)
    if cr.shape != (g.hc, g.wc) or cb.shape != cr.shape:
        raise ValueError("chroma planes %r/%r do not match luminance %r (expected %r)"
                         % (cr.shape, cb.shape, lum.shape, (g.hc, g.wc)))
    bufs = []
    for a in (lum, cr, cb):
        b = _lib.scratch(a.nbytes)
        b.upload(a, stream)
        bufs.append(b)
    coef = _lib.scratch(g.blocks_per_image * 128)
    _lib.check(lib.hic_planes_to_blocks(bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, 1, h, w, coef.ptr, stream))
    _lib.sync(stream)
    for b in bufs:
        b.free()
    return coef, g


def jpeg_decompression(d: model.CompressedImage) -> np.ndarray:
    """Dequantise, IDCT, +128, uint8 cast, chroma pyrUp, YCrCb -> RGB."""
    settings.check_supported()
    _lib.require_device()
    lib = _lib.load()
    coef, g = planes_to_device_blocks(d)
    y = _lib.scratch(g.h * g.w)
    cr = _lib.scratch(g.hc * g.wc)
    cb = _lib.scratch(g.hc * g.wc)
    rgb = _lib.scratch(g.out_h * g.out_w * 3)
    ties = _lib.scratch(g.blocks_per_image * _lib.TIE_RECORD_BYTES)
    stats = _lib.scratch(4 * _lib.TIE_STATS)
    _lib.check(lib.hic_dct_inverse(coef.ptr, 1, g.h, g.w, y.ptr, cr.ptr, cb.ptr, rgb.ptr, ties.ptr,
                                   g.blocks_per_image, stats.ptr, None))
    out = rgb.download(np.uint8, g.out_h * g.out_w * 3).reshape(g.out_h, g.out_w, 3)
    st = stats.download(np.uint32, _lib.TIE_STATS)
    LAST_INVERSE_STATS.update(flagged_blocks=int(st[0]), reevaluated=int(st[1]), changed=int(st[2]))
    for b in (coef, y, cr, cb, rgb, ties, stats):
        b.free()
    return out


def wavelet_compression(rgb_image: np.ndarray) -> model.CompressedImage:
    from hiccup_b200 import wavelet
    return wavelet.wavelet_compression(rgb_image)


def wavelet_decompression(channels: model.CompressedImage) -> np.ndarray:
    from hiccup_b200 import wavelet
    return wavelet.wavelet_decompression(channels)
