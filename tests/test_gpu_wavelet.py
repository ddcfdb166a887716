"""GPU parity of the wavelet denoise (row f4; publish/oe_map/S3_denoise.py:43-47) against
oracle/wavelet_oracle.py on the same inputs. The arithmetic is third-party and unversioned in the
reference (scikit-image + PyWavelets), so this pins the CUDA path to the written-down definition:
PARITY UNPINNED against the libraries themselves.

Tolerances: fp64 blocks within 1e-11 absolute (same arithmetic, fused multiply-adds and a different
summation order); fp32 blocks within 1e-5 absolute (north star) -- the error is the fp32 rounding of
the input and of the result. sigma / threshold agree to 1e-12 relative."""
import math

import numpy as np
import pytest
import torch

from oracle import wavelet_oracle as wo

pytestmark = pytest.mark.gpu


def _structured(shape, seed, noise=0.2):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    base = np.sin(yy / 17.0) * np.cos(xx / 23.0) + 0.3 * ((yy // 40 + xx // 40) % 2)
    return base + noise * rng.standard_normal(shape)


@pytest.mark.parametrize("shape", [(64, 64), (150, 120), (90, 150), (1, 9), (7, 3), (333, 517), (640, 1300), (33, 31)])
def test_wavelet_fp64_against_oracle(shape):
    from hiarch_b200 import ops
    x = _structured(shape, seed=shape[0] * 7 + shape[1])
    ref, sigma, thr, levels = wo.denoise_wavelet_db3_visushrink(x, return_parts=True)
    assert ops.wavelet_default_levels(*shape) == levels
    out, stats = ops.wavelet_denoise_db3(torch.from_numpy(x).cuda(), return_stats=True)
    stats = stats.cpu().numpy()
    # (a one-row block has only rounding noise, ~1e-17, in its finest diagonal band: absolute floor)
    assert abs(stats[0] - sigma) <= 1e-12 * sigma + 1e-15 and abs(stats[1] - thr) <= 1e-12 * thr + 1e-15
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-11


@pytest.mark.parametrize("levels", [1, 2, 3, 5])
def test_wavelet_explicit_levels(levels):
    from hiarch_b200 import ops
    x = _structured((200, 168), seed=levels)
    ref = wo.denoise_wavelet_db3_visushrink(x, levels=levels)
    out = ops.wavelet_denoise_db3(torch.from_numpy(x).cuda(), levels=levels)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-11


def test_wavelet_fp32_map_view():
    """The pipeline's form: an fp32 block that is a strided view into the genome-wide map."""
    from hiarch_b200 import ops
    full = _structured((400, 420), seed=11).astype(np.float32)
    dev = ops.alloc_map(400, "cuda", cols=420)
    dev.copy_(torch.from_numpy(full))
    blk = dev[37:290, 101:398]
    x = full[37:290, 101:398].astype(np.float64)
    ref = wo.denoise_wavelet_db3_visushrink(x)
    out = ops.wavelet_denoise_db3(blk)
    assert out.dtype == torch.float32 and tuple(out.shape) == x.shape
    assert np.abs(out.cpu().numpy().astype(np.float64) - ref).max() <= 1e-5
    assert torch.equal(dev.cpu(), torch.from_numpy(full))             # the input map is untouched


def test_wavelet_median_selection_with_zeros_and_ties():
    """The sigma estimate is an exact selection: blocks with many exact zeros (left out of the median),
    ties, an even and an odd number of nonzero coefficients."""
    from hiarch_b200 import ops
    rng = np.random.default_rng(5)
    for shape, keep in (((96, 96), 0.02), ((97, 95), 0.3), ((64, 80), 1.0)):
        x = np.round(rng.standard_normal(shape) * 4) / 4                # few distinct values -> ties
        x[rng.random(shape) > keep] = 0.0
        x[:8, :8] = rng.standard_normal((8, 8))
        ref, sigma, thr, _ = wo.denoise_wavelet_db3_visushrink(x, return_parts=True)
        out, stats = ops.wavelet_denoise_db3(torch.from_numpy(x).cuda(), return_stats=True)
        stats = stats.cpu().numpy()
        dd = wo.wavedec2(x, 1)[-1]["dd"]
        assert abs(stats[0] - sigma) <= 1e-12 * sigma, shape
        assert abs(stats[2] - np.count_nonzero(dd)) <= max(2, 1e-3 * dd.size)   # |dd| ~ 1e-17 either side of zero
        assert np.abs(out.cpu().numpy() - ref).max() <= 1e-11, shape


def test_wavelet_constant_block_is_nan_like_the_libraries():
    from hiarch_b200 import ops
    out, stats = ops.wavelet_denoise_db3(torch.zeros((40, 30), dtype=torch.float64, device="cuda"), return_stats=True)
    assert math.isnan(float(stats[0])) and float(stats[2]) == 0.0
    assert torch.isnan(out).all()


def test_wavelet_large_block_properties():
    """chr1 at 100 kb (2 490 bins): five levels. Size-independent checks: threshold 0 would be the
    identity (perfect reconstruction) -> here the residual is bounded by the threshold per coefficient;
    a 300-row stripe is compared with the oracle run on the whole block."""
    from hiarch_b200 import ops
    n = 2490
    x = _structured((n, n), seed=3, noise=0.3)
    out, stats = ops.wavelet_denoise_db3(torch.from_numpy(x).cuda(), return_stats=True)
    assert ops.wavelet_default_levels(n, n) == 5
    ref, sigma, thr, _ = wo.denoise_wavelet_db3_visushrink(x, return_parts=True)
    stats = stats.cpu().numpy()
    assert abs(stats[0] - sigma) <= 1e-12 * sigma
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-11
    assert np.abs(got - x).std() < 1.2 * 0.3          # what is removed is (at most) the noise


def test_main_denoise_with_the_wavelet_against_reference_run(golden):
    """main_denoise with the wavelet step on the device against the reference's own main_denoise run with
    oracle/wavelet_oracle.py standing in for skimage (tests/golden/gen_golden_normdis.py --wavelet db3)."""
    import pandas as pd
    from hiarch_b200 import oe_map, new_hic_class as nhc
    g = golden("gnormdis_db3")
    assert str(g["wavelet"]).startswith("db3")
    rows, idx = [], 0
    for name, n in zip((str(c) for c in g["pro.names"]), (int(x) for x in g["pro.lens"])):
        for b in range(n):
            rows.append((name, b * 100000, (b + 1) * 100000, idx))
            idx += 1
    win = nhc.GenomeIndex(pd.DataFrame(rows, columns=["chr", "start", "end", "index"]))
    m = nhc.ArrayHiCMtx(g["ode.dense"], row_window=win, col_window=win, mtx_type="all", has_neg=True)
    den = oe_map.main_denoise(m)
    assert den.wavelet_applied
    assert np.abs(den.matrix - g["denoised.dense"]).max() <= 2e-5
    ident = oe_map.main_denoise(m, wavelet="identity")
    assert not ident.wavelet_applied
    assert np.abs(ident.matrix - golden("gnormdis")["denoised.dense"]).max() <= 2e-5
