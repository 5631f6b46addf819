"""Host-side preparation of the wavelength-dependent line-spread function (array-valued ``Inst_R``).

``GenMod.genspec`` hands an array ``inst_R`` through unchanged (reference genmod.py:74-77) and the predictor then
smooths with ``smoothspec(wave, flux, sigma, outwave=outwave, smoothtype='lsf', fftsmooth=True)`` -- reference
minesweeper/smoothing.py:125-128 (dispatcher) and :412-514 (``smooth_lsf_fft``): the wavelength axis is warped by the
cumulative distribution of d(lambda) / sigma(lambda) so that the Gaussian has constant width in the warped coordinate
x in [0, 1], the spectrum is resampled on ``nx = 2**ceil(log2(pix_per_sigma / x_per_sigma))`` uniform x points, smoothed
there by FFT, and interpolated onto the observed wavelengths.

Everything up to the resampling map depends only on the model's wavelength grid and the star's sigma array: a Doppler
factor s multiplies the warped wavelengths and divides ``x_per_sigma``, nothing else.  ``prepare`` therefore runs once
per star on the host (NumPy f64, the reference's operation order) and ``Engine.set_lsf`` uploads the result through
``ms_set_lsf``; per-point work is in ``spec_lsf_kernel`` (csrc/spec_smooth.cu).
"""
import numpy as np


def prepare(wave, sigma, pix_per_sigma=2.0, max_velocity_kms=1000.0):
    """(nx, lam[nx], idx[nx] int32, frac[nx], x_per_sigma) for the model wavelengths ``wave`` and the LSF width
    ``sigma`` (same units, same length).  Raises if a Doppler shift of up to ``max_velocity_kms`` could change the
    power-of-two resampling length (the device keeps one map per star)."""
    wave = np.ascontiguousarray(wave, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    if wave.shape != sigma.shape or wave.ndim != 1:
        raise ValueError('LSF sigma array must have the length of the model wavelength grid (%d)' % len(wave))
    if not np.all(sigma > 0):
        raise ValueError('LSF sigma must be positive')
    dw = np.gradient(wave)                                   # smoothing.py:455
    cdf = np.cumsum(dw / sigma)
    cdf /= cdf.max()
    x_per_pixel = np.gradient(cdf)
    x_per_sigma = float(np.nanmedian(x_per_pixel / (dw / sigma)))
    length = pix_per_sigma / x_per_sigma
    nx = int(2.0 ** np.ceil(np.log2(length)))
    eps = max_velocity_kms / 299792.458
    for s in (1.0 - eps, 1.0 + eps):
        if int(2.0 ** np.ceil(np.log2(length * s))) != nx:
            raise ValueError('LSF array puts the resampling length (%.1f) within %.2f %% of a power of two: a Doppler '
                             'shift would change the grid; rescale sigma slightly' % (length, 100 * eps))
    x = np.linspace(0.0, 1.0, nx)
    lam = np.interp(x, cdf, wave)                            # warped wavelengths (smoothing.py:468-470)
    idx = np.clip(np.searchsorted(wave, lam, side='right') - 1, 0, len(wave) - 2)
    frac = (lam - wave[idx]) / (wave[idx + 1] - wave[idx])
    frac = np.clip(frac, 0.0, 1.0)                           # np.interp clamps at the ends
    return nx, lam, idx.astype(np.int32), frac, x_per_sigma
