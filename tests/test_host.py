"""Host-side logic of the product package that needs no GPU: the LSF preparation (row a14), sampler queue sizing, the
running-evidence recursion of the samples file."""
import os

import numpy as np
import pytest

from minesweeper_b200 import lsf
from minesweeper_b200.sampler import auto_queue, _running_evidence

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _finish_lsf_on_host(wave, flux, outwave, prep):
    """What spec_lsf_kernel does with the prepared map, in NumPy: gather-resample onto the warped grid, smooth with the
    Gaussian of x_per_sigma grid units by direct summation (the form the kernel uses), interpolate onto outwave."""
    nx, lam, idx, frac, xps = prep
    res = flux[idx] * (1.0 - frac) + flux[idx + 1] * frac
    sig_pix = xps * nx                                       # kernel width in grid points (the FFT's d = 1 / nx)
    half = int(np.ceil(8.5 * sig_pix))
    k = np.arange(-half, half + 1)
    taps = np.exp(-0.5 * (k / sig_pix) ** 2)
    taps /= taps.sum()
    ext = np.concatenate([res[-half:], res, res[:half]])     # the FFT convolution is circular
    conv = np.convolve(ext, taps, mode='valid')
    return np.interp(outwave, lam, conv)


def test_lsf_prepare_reproduces_the_reference_smoothing():
    """smoothing.py:412-514 (smooth_lsf_fft) through lsf.prepare + the kernel's arithmetic, against the reference's own
    output (tests/golden/smoothing.npz: lsf_out)."""
    g = np.load(os.path.join(GOLD, 'smoothing.npz'))
    w, f, ow = g['wave'], g['flux'], g['outwave']
    for (a, b), ref in zip(g['lsf_ab'], g['lsf_out']):
        prep = lsf.prepare(w, a + b * (w - w[0]))
        nx, lam, idx, frac, xps = prep
        assert nx & (nx - 1) == 0 and len(lam) == nx and idx.dtype == np.int32
        assert idx.min() >= 0 and idx.max() <= len(w) - 2 and frac.min() >= 0.0 and frac.max() <= 1.0
        assert 2.0 <= xps * nx < 4.0 + 1e-9                  # 2-4 grid points per sigma by construction of nx
        got = _finish_lsf_on_host(w, f, ow, prep)
        np.testing.assert_allclose(got, ref, rtol=2e-9, atol=2e-9)


def test_lsf_prepare_rejects_bad_input():
    w = np.linspace(5000.0, 5300.0, 2048)
    with pytest.raises(ValueError):
        lsf.prepare(w, np.ones(10))
    with pytest.raises(ValueError):
        lsf.prepare(w, np.zeros_like(w))
    # a width that puts the resampling length on a power of two: a Doppler shift would change the grid
    dw = w[1] - w[0]
    n = len(w)
    with pytest.raises(ValueError):
        lsf.prepare(w, np.full_like(w, 2.0 * dw * n / 4096.0 * (1.0 + 1e-4)))


def test_auto_queue_is_bounded_and_monotone():
    sizes = [1, 10, 500, 501, 5000, 5001, 40000, 40001, 150000, 150001, 10 ** 7]
    q = [auto_queue(n) for n in sizes]
    assert all(4 <= v <= 64 for v in q)
    assert q == sorted(q, reverse=True)                      # fewer walkers per star as the catalog grows
    assert auto_queue(100000) == 8 and auto_queue(20000) == 16 and auto_queue(1) == 64


def test_running_evidence_matches_direct_sums():
    rng = np.random.default_rng(3)
    lnl = np.sort(rng.normal(-50.0, 10.0, 400))
    logwt = lnl + np.log(np.diff(np.exp(-np.arange(401) / 150.0)) * -1.0)
    logz, h = _running_evidence(lnl, logwt)
    w = np.exp(logwt - logwt.max())
    for i in (0, 17, 399):
        z = np.log(w[:i + 1].sum()) + logwt.max()
        assert abs(logz[i] - z) < 1e-10
        p = np.exp(logwt[:i + 1] - z)
        assert abs(h[i] - ((p * lnl[:i + 1]).sum() - z)) < 1e-9
    # continuing from a carried state equals one pass
    lz2, h2 = _running_evidence(lnl[200:], logwt[200:], logz[199], h[199] + logz[199])
    np.testing.assert_allclose(lz2, logz[200:], rtol=0, atol=1e-10)
    np.testing.assert_allclose(h2, h[200:], rtol=0, atol=1e-9)
