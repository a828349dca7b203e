"""Host-only pieces of the wavelet mode at general settings (no GPU): the pyramid geometry the C ABI computes against
the shapes the oracle's transform produces, the order-statistic index of quantization.quality_threshold_value, and the
settings the CUDA path refuses."""
import numpy as np
import pytest

from oracle import hiccup_oracle as orc


@pytest.mark.parametrize("levels", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("shape", [(64, 96), (50, 38), (131, 70), (7, 5), (1, 1), (4320, 7680)])
def test_pyramid_geometry_matches_the_transform(levels, shape):
    from hiccup_b200 import _lib, wavelet
    g = _lib.wavelet_pyramid(shape[0], shape[1], levels)
    shapes = wavelet.band_shapes(g)
    assert g.levels == levels and g.n_bands == 3 * levels + 1 == len(shapes)
    if shape[0] * shape[1] <= 1 << 16:
        want = [b.shape for b in orc.wavelet_channel(np.zeros(shape, np.uint8), levels=levels)]
        assert shapes == want
    assert int(g.len) == sum(a * b for a, b in shapes)
    assert [int(g.band_off[i]) for i in range(g.n_bands)] == list(np.cumsum([0] + [a * b for a, b in shapes[:-1]]))
    # the three-level pyramid is the fused kernels' geometry
    if levels == 3:
        g3 = _lib.wavelet_geometry(shape[0], shape[1])
        assert int(g3.len) == int(g.len) and [int(g3.band_off[i]) for i in range(10)] == [int(g.band_off[i]) for i in range(10)]


def test_pyramid_argument_checks():
    from hiccup_b200 import _lib
    for levels in (0, 6, -1):
        with pytest.raises(_lib.HicError):
            _lib.wavelet_pyramid(64, 64, levels)
    with pytest.raises(_lib.HicError):
        _lib.wavelet_pyramid(0, 64, 3)


def test_quality_factor_index_is_the_references():
    """quantization.py:84-94: s[len(vals) - int(ceil(len(vals) * q))] of the ascending sort."""
    from hiccup_b200 import _lib, settings, wavelet
    g = _lib.wavelet_pyramid(50, 38, 3)
    n = int(g.len)
    saved = settings.WAVELET_QUALITY_FACTOR
    try:
        for q in (1, 0.9, 0.5, 0.25, 1e-3, 0.3333):
            settings.WAVELET_QUALITY_FACTOR = q
            p = wavelet._params(g)
            assert p.threshold_index == (-1 if q == 1 else n - int(np.ceil(n * q)))
            vals = np.random.default_rng(3).integers(-50, 50, n)
            if q != 1:
                assert orc.quality_threshold_value(vals, q) == np.sort(vals)[p.threshold_index]
        settings.WAVELET_QUALITY_FACTOR = 1e-9            # ceil() keeps at least one coefficient: the largest
        assert wavelet._params(g).threshold_index == n - 1
    finally:
        settings.WAVELET_QUALITY_FACTOR = saved


def test_settings_the_cuda_path_refuses():
    from hiccup_b200 import model, settings
    saved = (settings.WAVELET, settings.WAVELET_NUM_LEVELS, settings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
             settings.WAVELET_THRESHOLD, settings.WAVELET_QUALITY_FACTOR)
    try:
        settings.check_wavelet_supported()
        assert settings.wavelet_defaults()
        for wv in (model.Wavelet.COIF, model.Wavelet.SYM):
            settings.WAVELET = wv
            with pytest.raises(NotImplementedError):
                settings.check_wavelet_supported()
        settings.WAVELET = model.Wavelet.HAAR
        settings.check_wavelet_supported()
        for bad, exc in ((dict(WAVELET_NUM_LEVELS=0), NotImplementedError), (dict(WAVELET_NUM_LEVELS=6), NotImplementedError),
                         (dict(WAVELET_NUM_LEVELS=2.5), NotImplementedError), (dict(WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER=0), NotImplementedError),
                         (dict(WAVELET_QUALITY_FACTOR=0), ValueError), (dict(WAVELET_QUALITY_FACTOR=1.5), ValueError),
                         (dict(WAVELET_THRESHOLD=-1), ValueError)):
            for k, v in bad.items():
                old = getattr(settings, k)
                setattr(settings, k, v)
                with pytest.raises(exc):
                    settings.check_wavelet_supported()
                setattr(settings, k, old)
        settings.WAVELET_NUM_LEVELS = 4
        assert not settings.wavelet_defaults()
    finally:
        (settings.WAVELET, settings.WAVELET_NUM_LEVELS, settings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
         settings.WAVELET_THRESHOLD, settings.WAVELET_QUALITY_FACTOR) = saved
