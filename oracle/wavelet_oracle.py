"""CPU restatement of the wavelet step of the reference's denoise. TEST INFRASTRUCTURE ONLY: nothing
under hiarch_b200/ may import this file (tests, smoke() and the golden generators do).

PARITY UNPINNED. The reference calls

    skimage.restoration.denoise_wavelet(dense_arr, wavelet='db3', method="VisuShrink")
                                                   (publish/oe_map/S3_denoise.py:43-47)

and the arithmetic lives in scikit-image + PyWavelets, which are absent from /root/reference, not
installed here and not version-pinned anywhere in the reference (no requirements file). This file
writes the definition down from the two libraries' published algorithms:

  PyWavelets  pywt.wavedecn / pywt.waverecn / pywt.dwt / pywt.idwt / pywt.threshold, signal
              extension mode 'symmetric' (half-sample symmetric: ... x1 x0 | x0 x1 ... xn-1 | xn-1 xn-2 ...):
      dwt :  cA[k] = sum_j dec_lo[j] * x_ext[2k + 1 - j],  k = 0 .. floor((N + F - 1) / 2) - 1
      idwt:  x[n]  = sum_k cA[k] * rec_lo[n - 2k + F - 2] + cD[k] * rec_hi[n - 2k + F - 2],
                     n = 0 .. 2 len(cA) - F + 1
      wavedecn: per level, dwt along every axis of the running approximation ('aa' 'ad' 'da' 'dd',
              first letter = axis 0); waverecn: trims the running approximation by one along an axis
              when it is one longer than the level's detail arrays (odd lengths), then idwt per axis.
      dwt_max_level(n, F) = floor(log2(n / (F - 1))) (0 when n < F - 1).
  scikit-image (>= 0.16 behaviour) restoration/_denoise.py:
      denoise_wavelet -> _wavelet_threshold:
          levels    = max(dwtn_max_level(shape, wavelet) - 3, 1)
          sigma     = median(|dd_finest[dd_finest != 0]|) / norm.ppf(0.75)       (_sigma_est_dwt)
          threshold = sigma * sqrt(2 ln(image.size))                             (VisuShrink, _universal_thresh)
          details   -> pywt.threshold(d, threshold, mode='soft') = d * max(1 - threshold/|d|, 0)
          out       = waverecn(...)[original extent]
      floating-point input is neither rescaled nor clipped (older releases, <= 0.15, clipped the
      result to [-1, 1] / [0, 1]; `clip_legacy=True` restates that variant).

The db3 filter is the mathematical one (closed form below; tests/test_wavelet_oracle.py re-derives it
by spectral factorisation); PyWavelets' table agrees with it to ~1e-12.

What IS pinned by tests: the filter (orthonormality, 3 vanishing moments, the spectral
factorisation), perfect reconstruction of dwt/idwt and wavedec2/waverec2 for every length parity,
the output lengths PyWavelets documents, the soft threshold, and hand-computed known answers.
"""
import math

import numpy as np

_S10 = math.sqrt(10.0)
_S5 = math.sqrt(5.0 + 2.0 * _S10)
_D = 16.0 * math.sqrt(2.0)
# scaling filter h0..h5 of the Daubechies wavelet with 3 vanishing moments (= pywt rec_lo)
DB3_REC_LO = np.array([(1 + _S10 + _S5) / _D, (5 + _S10 + 3 * _S5) / _D, (10 - 2 * _S10 + 2 * _S5) / _D,
                       (10 - 2 * _S10 - 2 * _S5) / _D, (5 + _S10 - 3 * _S5) / _D, (1 + _S10 - _S5) / _D])
DB3_DEC_LO = DB3_REC_LO[::-1].copy()
# pywt: dec_hi[j] = (-1)^(j+1) dec_lo[F-1-j]; rec_hi = dec_hi reversed
DB3_DEC_HI = np.array([(-1.0) ** (j + 1) * DB3_DEC_LO[5 - j] for j in range(6)])
DB3_REC_HI = DB3_DEC_HI[::-1].copy()
F = 6
NORM_PPF_075 = 0.6744897501960817      # scipy.stats.norm.ppf(0.75)


def dwt_len(n):
    """pywt.dwt_coeff_len(n, F, 'symmetric')."""
    return (n + F - 1) // 2


def dwt_max_level(n):
    """pywt.dwt_max_level(n, F)."""
    if n < F - 1:
        return 0
    return int(math.floor(math.log2(n / (F - 1))))


def default_levels(shape):
    """skimage _wavelet_threshold: skip the three coarsest scales, at least one level."""
    return max(min(dwt_max_level(s) for s in shape) - 3, 1)


def _dwt_last(x, filt):
    """One filter of pywt.dwt(mode='symmetric') along the last axis."""
    n = x.shape[-1]
    L = dwt_len(n)
    pad = [(0, 0)] * (x.ndim - 1) + [(F - 1, F - 1)]
    ext = np.pad(x, pad, mode="symmetric")                      # ext[p] = x[p - (F-1)]
    out = np.zeros(x.shape[:-1] + (L,), dtype=np.float64)
    k = np.arange(L)
    for j in range(F):
        out += filt[j] * ext[..., 2 * k + 1 - j + (F - 1)]
    return out


def dwt1(x, axis):
    """(cA, cD) = pywt.dwt(x, 'db3', mode='symmetric', axis=axis)."""
    xm = np.moveaxis(np.asarray(x, dtype=np.float64), axis, -1)
    return (np.moveaxis(_dwt_last(xm, DB3_DEC_LO), -1, axis),
            np.moveaxis(_dwt_last(xm, DB3_DEC_HI), -1, axis))


def idwt1(ca, cd, axis):
    """pywt.idwt(cA, cD, 'db3', mode='symmetric', axis=axis): length 2 len(cA) - F + 2."""
    a = np.moveaxis(np.asarray(ca, dtype=np.float64), axis, -1)
    d = np.moveaxis(np.asarray(cd, dtype=np.float64), axis, -1)
    L = a.shape[-1]
    n_out = 2 * L - F + 2
    out = np.zeros(a.shape[:-1] + (max(n_out, 0),), dtype=np.float64)
    n = np.arange(n_out)
    for k in range(L):
        t = n - 2 * k + F - 2
        ok = (t >= 0) & (t < F)
        if not ok.any():
            continue
        out[..., n[ok]] += a[..., k:k + 1] * DB3_REC_LO[t[ok]] + d[..., k:k + 1] * DB3_REC_HI[t[ok]]
    return np.moveaxis(out, -1, axis)


def dwt2(a):
    """One level of pywt.dwtn on a 2-D array: {'aa','ad','da','dd'}, first letter = axis 0."""
    lo0, hi0 = dwt1(a, 0)
    aa, ad = dwt1(lo0, 1)
    da, dd = dwt1(hi0, 1)
    return {"aa": aa, "ad": ad, "da": da, "dd": dd}


def idwt2(c):
    lo0 = idwt1(c["aa"], c["ad"], 1)
    hi0 = idwt1(c["da"], c["dd"], 1)
    return idwt1(lo0, hi0, 0)


def wavedec2(img, levels):
    """pywt.wavedecn(img, 'db3', level=levels): [aa_coarsest, {ad,da,dd}_coarsest, ..., {..}_finest]."""
    a = np.asarray(img, dtype=np.float64)
    details = []
    for _ in range(levels):
        c = dwt2(a)
        a = c.pop("aa")
        details.append(c)
    return [a] + details[::-1]


def waverec2(coeffs):
    """pywt.waverecn: the running approximation is trimmed to the detail shape before each level."""
    a = coeffs[0]
    for d in coeffs[1:]:
        shp = d["dd"].shape
        a = a[:shp[0], :shp[1]]
        a = idwt2({"aa": a, **d})
    return a


def soft_threshold(d, value):
    """pywt.threshold(d, value, mode='soft')."""
    d = np.asarray(d, dtype=np.float64)
    mag = np.abs(d)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = 1.0 - value / mag
        f = np.clip(f, 0.0, None)          # nan stays nan, like ndarray.clip
        out = d * f
    if not np.isnan(value):
        out = np.where(mag == 0.0, 0.0, out)   # pywt zeroes the 0/0 entries
    return out


def sigma_est(dd):
    """skimage _sigma_est_dwt: MAD of the nonzero finest diagonal coefficients (nan when none)."""
    nz = np.abs(dd[dd != 0])
    if nz.size == 0:
        return float("nan")
    return float(np.median(nz)) / NORM_PPF_075


def denoise_wavelet_db3_visushrink(img, levels=None, clip_legacy=False, return_parts=False):
    """skimage.restoration.denoise_wavelet(img, wavelet='db3', method='VisuShrink') for a 2-D float
    array (soft threshold, sigma estimated, levels = default_levels)."""
    img = np.asarray(img, dtype=np.float64)
    if img.ndim != 2:
        raise ValueError("2-D array expected")
    if levels is None:
        levels = default_levels(img.shape)
    coeffs = wavedec2(img, levels)
    sigma = sigma_est(coeffs[-1]["dd"])
    thr = sigma * math.sqrt(2.0 * math.log(img.size)) if img.size > 0 else float("nan")
    den = [coeffs[0]] + [{k: soft_threshold(v, thr) for k, v in lvl.items()} for lvl in coeffs[1:]]
    out = waverec2(den)[:img.shape[0], :img.shape[1]]
    if clip_legacy:
        lo, hi = ((-1.0, 1.0) if img.min() < 0 else (0.0, 1.0))
        out = np.clip(out, lo, hi)
    if return_parts:
        return out, sigma, thr, levels
    return out
