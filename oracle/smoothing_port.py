"""ORACLE (test infrastructure only) -- NumPy f64 restatement of the two
smoothing modes that sit on the likelihood path.

Follows reference minesweeper/smoothing.py:
  constants            :14-15   (ckms = 2.998e5, sigma_to_fwhm = 2.355)
  dispatcher           :18-158  (smoothtype 'R' and 'vsini', fftsmooth=True)
  smooth_vel_fft       :202-240
  smooth_vsini_fft     :242-264
  smooth_fft           :517-537
  smooth_fft_vsini     :539-558
  mask_wave            :560-576
  resample_wave        :579-598
Pinned against the reference module itself by tests/golden/smoothing.npz
(written by tests/golden/make_golden.py) -- see tests/test_oracle.py.
"""
import numpy as np
from scipy.special import j1

CKMS = 2.998e5          # reference smoothing.py:14 (deliberately not 299792.458)
NSIGMA_PAD = 20.0       # reference smoothing.py:561


def pow2_ceil(n):
    """Smallest power of two >= n (reference smoothing.py:587)."""
    return int(2.0 ** np.ceil(np.log2(n)))


def log_resample(wave, spec):
    """Linear interpolation onto a log-uniform grid with a power-of-two
    number of points spanning [wave.min(), wave.max()] (smoothing.py:579-598)."""
    n = pow2_ceil(len(wave))
    grid = np.exp(np.linspace(np.log(wave.min()), np.log(wave.max()), n))
    return grid, np.interp(grid, wave, spec)


def padded_window(wave, rsigma, outwave):
    """Boolean mask of input pixels inside the output range padded by
    20 / rsigma relative on each side, strict inequalities
    (smoothing.py:560-576 with linear=False).  outwave=None keeps every pixel
    (limits become 0 and +inf)."""
    if outwave is None:
        lo, hi = 0.0, np.inf
    else:
        lo, hi = outwave.min(), outwave.max()
    lo = lo * (1.0 - NSIGMA_PAD / rsigma)
    hi = hi * (1.0 + NSIGMA_PAD / rsigma)
    return (wave > lo) & (wave < hi)


def gaussian_transfer(n, dv, sigma):
    """Analytic Fourier transform of a unit-area Gaussian sampled at the
    rfft frequencies of an n-point grid of spacing dv (smoothing.py:527-529)."""
    f = np.fft.rfftfreq(n, d=dv)
    return np.exp(-2.0 * np.pi ** 2 * sigma ** 2 * f ** 2)


def rotation_transfer(n, dv, vsini):
    """Transfer function of the rotational profile used by the reference
    (smoothing.py:541-548); DC forced to one."""
    f = np.fft.rfftfreq(n, d=dv)
    f[0] = 0.01
    u = 2.0 * np.pi * vsini * f
    t = j1(u) / u - 3.0 * np.cos(u) / (2.0 * u ** 2) + 3.0 * np.sin(u) / (2.0 * u ** 3)
    t[0] = 1.0
    return t


def _fft_filter(wave, spec, outwave, sigma, transfer):
    """Common scaffold of smooth_vel_fft / smooth_vsini_fft
    (smoothing.py:218-240, 244-264)."""
    if sigma <= 0:
        return np.interp(outwave, wave, spec)
    grid, s = log_resample(wave, spec)
    step = np.diff(np.log(grid))
    assert step.max() / step.min() < 1.05
    dv = CKMS * np.median(step)
    conv = np.fft.irfft(np.fft.rfft(s) * transfer(len(s), dv, sigma))
    return np.interp(outwave, grid, conv)


def smooth_R(wave, spec, rsigma, outwave, inres=None):
    """smoothspec(..., smoothtype='R', fftsmooth=True, inres=inres)
    (smoothing.py:102-114, 131-134, 202-240).  ``rsigma`` and ``inres`` are
    lambda/sigma_lambda; the kernel width is their quadrature difference in
    km/s (NaN when the target is sharper than the input, which the reference
    lets propagate)."""
    sigma_out = CKMS / rsigma
    sigma_in = 0.0 if inres is None else CKMS / inres
    keep = padded_window(wave, rsigma, outwave)
    w, s = wave[keep], spec[keep]
    if outwave is None:
        outwave = wave
    with np.errstate(invalid='ignore'):
        sigma = np.sqrt(sigma_out ** 2 - sigma_in ** 2)
    return _fft_filter(w, s, outwave, sigma, gaussian_transfer)


def smooth_vsini(wave, spec, vsini, outwave=None, inres=0.0):
    """smoothspec(..., smoothtype='vsini', fftsmooth=True)
    (smoothing.py:92-99, 242-264)."""
    rsigma = CKMS / vsini
    keep = padded_window(wave, rsigma, outwave)
    w, s = wave[keep], spec[keep]
    if outwave is None:
        outwave = wave
    sigma = np.sqrt(vsini ** 2 - inres ** 2)
    return _fft_filter(w, s, outwave, sigma, rotation_transfer)


def smooth_lsf(wave, spec, outwave, sigma, pix_per_sigma=2.0):
    """smoothspec(..., smoothtype='lsf', fftsmooth=True) with an array ``sigma`` (same units and length
    as ``wave``): wavelength-dependent Gaussian line-spread function (smoothing.py:125-128, 412-514).

    The wavelength axis is warped by the cumulative distribution of d(lambda) / sigma(lambda) so that the
    kernel has constant width in the warped coordinate x in [0, 1]; the spectrum is resampled on
    ``nx = 2**ceil(log2(pix_per_sigma / x_per_sigma))`` uniform x points, smoothed there by FFT with a
    Gaussian of ``x_per_sigma`` (the median over pixels of dx / (d lambda / sigma)), and interpolated onto
    ``outwave``.  The reference's mask for this mode pads by 20 x 100 wavelength units (smoothing.py:
    125-132), i.e. keeps the whole model spectrum for any realistic window, so no masking is restated.
    Row a14 of SURVEY.md section 8; the CUDA path (``spec_lsf_kernel``) is tested against this restatement and
    the reference's golden vectors."""
    wave = np.asarray(wave, dtype=np.float64)
    spec = np.asarray(spec, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    dw = np.gradient(wave)
    cdf = np.cumsum(dw / sigma)
    cdf = cdf / cdf.max()
    x_per_sigma = np.nanmedian(np.gradient(cdf) / (dw / sigma))
    nx = int(pow2_ceil(pix_per_sigma / x_per_sigma))
    x = np.linspace(0.0, 1.0, nx)
    lam = np.interp(x, cdf, wave)
    resampled = np.interp(lam, wave, spec)
    f = np.fft.rfftfreq(nx, d=1.0 / nx)
    conv = np.fft.irfft(np.fft.rfft(resampled) * np.exp(-2.0 * np.pi ** 2 * x_per_sigma ** 2 * f ** 2))
    return np.interp(outwave, lam, conv)
