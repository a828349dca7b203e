"""Stand-in for the slice of PyWavelets the reference touches.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: PyWavelets is a third-party dependency of the reference (setup.py:26, unpinned
version) that is NOT installed in this image and not in /opt/wheelhouse, so nothing here could be
checked against the real package.  The reference's own tests pin only sub-band shapes
(transformtest.py:163-168, 50 -> 25 -> 13) and perfect reconstruction (transformtest.py:170-175);
both hold for this restatement (tests/test_oracle_wavelet.py).

What is restated (call sites: reference hiccup/transform.py:200 `pywt.wavedec2`, :224
`pywt.waverec2`), following PyWavelets' published algorithm:

  * `db1` / `haar` decomposition filters  dec_lo = [c, c], dec_hi = [-c, c],
    reconstruction filters               rec_lo = [c, c], rec_hi = [c, -c],  c = 1/sqrt(2)
    (pywt stores 7.071067811865475244e-01, which rounds to the double 0x1.6a09e667f3bcdp-1).
  * mode "symmetric" (pywt default): half-sample symmetric extension, x[N] = x[N-1].
  * `dwt` downsampling convolution: out[k] = sum_j filt[j] * x[2k+1-j], accumulated in j order
    starting from 0.0 (pywt/_extensions/c/convolution.template.c), so
        cA[k] = (c*x[2k+1]) + (c*x[2k]),   cD[k] = (-c*x[2k+1]) + (c*x[2k])   in float64.
  * `dwtn`: axis 0 first, then axis 1; `dwt2` returns cA='aa', (cH='da', cV='ad', cD='dd')
    where the first letter is axis 0.
  * `wavedec2` returns [cA_n, (cH_n, cV_n, cD_n), ..., (cH_1, cV_1, cD_1)].
  * `idwt`: upsampling convolution accumulating the low-pass branch first, then the high-pass:
        x[2k] = (c*cA[k]) + (c*cD[k]),   x[2k+1] = (c*cA[k]) + (-c*cD[k]);
    `idwtn` undoes the LAST axis first; `waverec2` trims one row/column when the running
    approximation is one larger than the next detail band.
"""
import numpy as np

C = np.float64(0.7071067811865475244008443621048490392848359376884740365883398)


class Wavelet:
    def __init__(self, name):
        if name not in ("db1", "haar"):
            raise ValueError("pywt stand-in only restates db1/haar, not %r" % (name,))
        self.name = name


def _as_name(wavelet):
    return wavelet.name if isinstance(wavelet, Wavelet) else wavelet


def _check(wavelet):
    if _as_name(wavelet) not in ("db1", "haar"):
        raise ValueError("pywt stand-in only restates db1/haar, not %r" % (wavelet,))


def _dwt_axis(x, axis):
    x = np.asarray(x, dtype=np.float64)
    x = np.moveaxis(x, axis, -1)
    n = x.shape[-1]
    if n % 2:
        x = np.concatenate([x, x[..., -1:]], axis=-1)     # symmetric extension
    even = x[..., 0::2]
    odd = x[..., 1::2]
    ca = (C * odd) + (C * even)
    cd = ((-C) * odd) + (C * even)
    return np.moveaxis(ca, -1, axis), np.moveaxis(cd, -1, axis)


def _idwt_axis(ca, cd, axis):
    ca = np.moveaxis(np.asarray(ca, dtype=np.float64), axis, -1)
    cd = np.moveaxis(np.asarray(cd, dtype=np.float64), axis, -1)
    assert ca.shape == cd.shape
    out = np.empty(ca.shape[:-1] + (2 * ca.shape[-1],), dtype=np.float64)
    out[..., 0::2] = (C * ca) + (C * cd)
    out[..., 1::2] = (C * ca) + ((-C) * cd)
    return np.moveaxis(out, -1, axis)


def dwt2(data, wavelet, mode="symmetric"):
    _check(wavelet)
    a, d = _dwt_axis(data, 0)
    aa, ad = _dwt_axis(a, 1)
    da, dd = _dwt_axis(d, 1)
    return aa, (da, ad, dd)


def idwt2(coeffs, wavelet, mode="symmetric"):
    _check(wavelet)
    aa, (da, ad, dd) = coeffs
    a = _idwt_axis(aa, ad, 1)
    d = _idwt_axis(da, dd, 1)
    return _idwt_axis(a, d, 0)


def wavedec2(data, wavelet, mode="symmetric", level=None):
    _check(wavelet)
    a = np.asarray(data)
    out = []
    for _ in range(level):
        a, ds = dwt2(a, wavelet, mode)
        out.append(ds)
    out.append(a)
    out.reverse()
    return out


def waverec2(coeffs, wavelet, mode="symmetric"):
    _check(wavelet)
    a = np.asarray(coeffs[0], dtype=np.float64)
    for d in coeffs[1:]:
        d = tuple(np.asarray(x, dtype=np.float64) for x in d)
        if a.shape[-2] == d[0].shape[-2] + 1:
            a = a[:-1, :]
        if a.shape[-1] == d[0].shape[-1] + 1:
            a = a[:, :-1]
        a = idwt2((a, d), wavelet, mode)
    return a
