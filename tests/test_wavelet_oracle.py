"""oracle/wavelet_oracle.py (the written-down db3 / VisuShrink definition, parity unpinned: scikit-image
and PyWavelets are absent and unversioned in the reference) against what CAN be pinned without the two
libraries: the filter's mathematical definition, the transform's documented lengths and perfect
reconstruction, the threshold rule, hand-computed known answers."""
import math

import numpy as np
import pytest

from oracle import wavelet_oracle as wo


def daubechies_by_spectral_factorisation(p):
    """Minimum-phase scaling filter with p vanishing moments: zeros of P(y) = sum_k C(p-1+k, k) y^k mapped to
    z through y = (2 - z - 1/z) / 4, the roots inside the unit circle kept, times (1 + z)^p."""
    coeffs = [math.comb(p - 1 + k, k) for k in range(p)]                # P(y), ascending
    roots_y = np.roots(coeffs[::-1])
    zs = []
    for y in roots_y:
        # z^2 - (2 - 4y) z + 1 = 0
        b = 2.0 - 4.0 * y
        disc = np.sqrt(b * b - 4.0 + 0j)
        z1, z2 = (b + disc) / 2.0, (b - disc) / 2.0
        zs.append(z1 if abs(z1) < 1 else z2)
    poly = np.array([1.0 + 0j])
    for _ in range(p):
        poly = np.convolve(poly, [1.0, 1.0])
    for z in zs:
        poly = np.convolve(poly, [1.0, -z])
    h = np.real(poly)
    return h * (math.sqrt(2.0) / h.sum())


def test_db3_filter_is_the_daubechies_filter():
    h = daubechies_by_spectral_factorisation(3)
    assert np.allclose(h, wo.DB3_REC_LO, atol=1e-13, rtol=0)
    # the digits every wavelet table prints for db3
    assert np.allclose(wo.DB3_REC_LO, [0.3326705530, 0.8068915093, 0.4598775021, -0.1350110200, -0.0854412739,
                                       0.0352262919], atol=1e-9, rtol=0)
    assert np.array_equal(wo.DB3_DEC_LO, wo.DB3_REC_LO[::-1])
    assert np.array_equal(wo.DB3_REC_HI, wo.DB3_DEC_HI[::-1])
    assert wo.DB3_DEC_HI[0] < 0 and wo.DB3_DEC_HI[1] > 0                 # PyWavelets' sign convention


def test_db3_filter_properties():
    lo, hi = wo.DB3_DEC_LO, wo.DB3_DEC_HI
    assert abs(lo.sum() - math.sqrt(2.0)) < 1e-15 and abs(hi.sum()) < 1e-15
    for shift in (0, 2, 4):
        dot = float(np.dot(lo[shift:], lo[:6 - shift]))
        assert abs(dot - (1.0 if shift == 0 else 0.0)) < 1e-15
        assert abs(float(np.dot(lo[shift:], hi[:6 - shift]))) < 1e-15
        assert abs(float(np.dot(hi[shift:], lo[:6 - shift]))) < 1e-15
    j = np.arange(6.0)
    for m in range(3):                                                   # three vanishing moments
        assert abs(float(np.dot(hi, j ** m))) < 1e-13


def test_lengths_and_levels_follow_pywavelets():
    assert [wo.dwt_len(n) for n in (1, 2, 5, 6, 7, 100, 101)] == [3, 3, 5, 5, 6, 52, 53]
    assert [wo.dwt_max_level(n) for n in (1, 4, 5, 9, 10, 19, 20, 2490)] == [0, 0, 0, 0, 1, 1, 2, 8]
    assert wo.default_levels((2490, 2490)) == 5 and wo.default_levels((150, 120)) == 1
    assert wo.default_levels((2490, 3)) == 1 and wo.default_levels((640, 1300)) == 4


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 17, 100, 101])
def test_dwt_idwt_perfect_reconstruction(n):
    x = np.random.default_rng(n).standard_normal(n)
    ca, cd = wo.dwt1(x, 0)
    assert len(ca) == len(cd) == wo.dwt_len(n)
    rec = wo.idwt1(ca, cd, 0)
    assert len(rec) == 2 * len(ca) - 4 and len(rec) in (n, n + 1) or n < 4
    assert np.abs(rec[:n] - x).max() < 1e-14


def test_dwt_known_answers():
    # a constant has no detail, and its approximation is the constant times sqrt(2) (symmetric extension keeps it flat)
    ca, cd = wo.dwt1(np.full(11, 3.0), 0)
    assert np.allclose(ca, 3.0 * math.sqrt(2.0), atol=1e-14) and np.abs(cd).max() < 1e-14
    # quadratics are annihilated by the detail filter away from the borders
    t = np.arange(40.0)
    _, cd = wo.dwt1(0.5 * t * t - 3 * t + 1, 0)
    assert np.abs(cd[3:-3]).max() < 1e-10 and np.abs(cd[:2]).max() > 1e-3
    # a unit impulse in the interior reproduces the (reversed, decimated) filters
    x = np.zeros(20); x[10] = 1.0
    ca, cd = wo.dwt1(x, 0)
    # cA[k] = dec_lo[2k + 1 - 10]
    for k in range(len(ca)):
        j = 2 * k + 1 - 10
        assert ca[k] == (wo.DB3_DEC_LO[j] if 0 <= j < 6 else 0.0)
        assert cd[k] == (wo.DB3_DEC_HI[j] if 0 <= j < 6 else 0.0)


@pytest.mark.parametrize("shape", [(17, 23), (64, 64), (101, 7), (1, 9), (30, 255)])
@pytest.mark.parametrize("levels", [1, 2, 3])
def test_wavedec2_waverec2_roundtrip(shape, levels):
    x = np.random.default_rng(7).standard_normal(shape)
    c = wo.wavedec2(x, levels)
    assert len(c) == levels + 1
    r, q = shape
    for lvl in c[:0:-1]:                                                # finest first
        r, q = wo.dwt_len(r), wo.dwt_len(q)
        assert {k: v.shape for k, v in lvl.items()} == {"ad": (r, q), "da": (r, q), "dd": (r, q)}
    assert c[0].shape == (r, q)
    rec = wo.waverec2(c)
    assert np.abs(rec[:shape[0], :shape[1]] - x).max() < 1e-13
    # orthonormal + symmetric extension: the energy can only grow by the mirrored border samples
    e = sum(float((v * v).sum()) for lvl in c[1:] for v in lvl.values()) + float((c[0] ** 2).sum())
    assert e >= float((x * x).sum()) * (1 - 1e-12)


def test_soft_threshold():
    d = np.array([-3.0, -1.0, -0.5, 0.0, 0.25, 1.0, 2.5])
    assert np.allclose(wo.soft_threshold(d, 1.0), [-2.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.5], atol=1e-15)
    assert np.isnan(wo.soft_threshold(d, float("nan"))).all()


def test_sigma_estimate_and_visushrink_threshold():
    rng = np.random.default_rng(3)
    noise = 0.37 * rng.standard_normal((400, 300))
    out, sigma, thr, levels = wo.denoise_wavelet_db3_visushrink(noise, return_parts=True)
    assert abs(sigma / 0.37 - 1.0) < 0.02                                # MAD / 0.6745 recovers the noise level
    assert thr == sigma * math.sqrt(2.0 * math.log(noise.size)) and levels == wo.default_levels(noise.shape)
    assert out.shape == noise.shape and out.std() < 0.3 * noise.std()   # only the coarse approximation keeps noise
    # exact zeros are left out of the median
    dd = np.array([[0.0, 2.0, 0.0], [-4.0, 0.0, 6.0]])
    assert wo.sigma_est(dd) == 4.0 / wo.NORM_PPF_075
    assert math.isnan(wo.sigma_est(np.zeros((3, 3))))
    # a smooth image plus small noise comes back close to the smooth image
    yy, xx = np.mgrid[0:200, 0:240]
    smooth = np.sin(yy / 31.0) * np.cos(xx / 47.0)
    den = wo.denoise_wavelet_db3_visushrink(smooth + 0.05 * rng.standard_normal(smooth.shape))
    assert np.abs(den - smooth).mean() < 0.02


def test_constant_block_follows_the_libraries_into_nan():
    """No nonzero finest-diagonal coefficient: numpy's median of nothing is nan, the threshold is nan and
    pywt.threshold propagates it; the restatement (and the device path) do the same rather than invent a rule."""
    out = wo.denoise_wavelet_db3_visushrink(np.zeros((12, 12)))
    assert np.isnan(out).all()


def test_legacy_clip_variant():
    x = np.random.default_rng(5).standard_normal((40, 40)) * 3
    a = wo.denoise_wavelet_db3_visushrink(x)
    b = wo.denoise_wavelet_db3_visushrink(x, clip_legacy=True)
    assert np.array_equal(b, np.clip(a, -1, 1))
