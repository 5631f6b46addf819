"""ORACLE (test infrastructure only) -- NumPy f64 restatement of the three
entry points of the external ``Payne`` package (github.com/pacargile/ThePayne)
that the reference calls on the likelihood path.

******************************  PARITY UNPINNED  ******************************
``Payne`` is not vendored under /root/reference, is not pinned in the
reference's setup.py:29 or README.md:38-46, is not installed and cannot be
downloaded here.  What follows restates its *published* algorithm as recalled
in SURVEY.md Appendix B and anchors on the reference's own call sites:

  FastPayneSEDPredict(usebands, nnpath)      genmod.py:39-44   (.filternames :44)
  FastPayneSEDPredict.sed(logt, logg, feh, afe, logl, av, dist, rv)   genmod.py:116-129
      magnitude formula corroborated by the legacy comment at genmod.py:133-140
  PayneSpecPredict(nnpath[, Cnnpath])        genmod.py:24-34
  PayneSpecPredict.getspec(Teff, logg, feh, afe, rad_vel, rot_vel, vmic,
                           inst_R, outwave)  genmod.py:80-84
  Payne.utils.quantiles.quantile             advancedpriors.py:32, 666
There are no golden vectors for these anywhere in the reference, so the ANN
arithmetic below defines, rather than reproduces, the behaviour the CUDA path
is tested against.  Architecture, activation and scalers are read from the
weight container so the port follows real files once they exist.
*******************************************************************************

Weight containers are dicts of arrays (or ``.npz`` paths) as produced by
``minesweeper_b200.synth.make_phot_net`` / ``make_spec_net``.
"""
import numpy as np

from . import smoothing_port as sm

C_KMS = 299792.458      # Payne shifts wavelengths with scipy.constants.c / 1000


def _load(container):
    if isinstance(container, (str, bytes)) or hasattr(container, '__fspath__'):
        with np.load(container, allow_pickle=False) as z:
            return {k: z[k] for k in z.files}
    return container


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


class FastPayneSEDPredict(object):
    """Stacked per-band MLPs predicting bolometric corrections."""

    def __init__(self, usebands=None, nnpath=None):
        net = _load(nnpath)
        names = [str(b) for b in net['bands']]
        if usebands is None:
            pick = list(range(len(names)))
        else:
            pick = [names.index(str(b)) for b in usebands if str(b) in names]
        self.filternames = [names[i] for i in pick]
        f8 = np.float64
        self.w1 = net['w1'][pick].astype(f8)
        self.b1 = net['b1'][pick].astype(f8)
        self.w2 = net['w2'][pick].astype(f8)
        self.b2 = net['b2'][pick].astype(f8)
        self.w3 = net['w3'][pick].astype(f8)
        self.b3 = net['b3'][pick].astype(f8)
        self.xmin = net['xmin'].astype(f8)
        self.xmax = net['xmax'].astype(f8)
        self.d_in = self.w1.shape[2]

    def bolcorr(self, x):
        """BC for every band; x = (Teff, logg, feh, afe, av[, rv])."""
        xs = (np.asarray(x, dtype=np.float64)[:self.d_in] - self.xmin) / (self.xmax - self.xmin) - 0.5
        a1 = _sigmoid(np.einsum('bhd,d->bh', self.w1, xs) + self.b1)
        a2 = _sigmoid(np.einsum('bhk,bk->bh', self.w2, a1) + self.b2)
        return np.einsum('bok,bk->bo', self.w3, a2)[:, 0] + self.b3[:, 0]

    def bolcorr_batch(self, X):
        """Vectorised ``bolcorr`` for X[B, d_in] (test helper; same arithmetic)."""
        xs = (np.asarray(X, dtype=np.float64)[:, :self.d_in] - self.xmin) / (self.xmax - self.xmin) - 0.5
        out = np.empty((len(xs), len(self.w1)))
        for b in range(len(self.w1)):                  # one BLAS product per band and layer
            a1 = _sigmoid(xs @ self.w1[b].T + self.b1[b])
            a2 = _sigmoid(a1 @ self.w2[b].T + self.b2[b])
            out[:, b] = a2 @ self.w3[b, 0] + self.b3[b, 0]
        return out

    def sed_batch(self, photpars):
        """photpars[B,7] = (Teff, logg, feh, afe, logR, Dist, Av) -> mags[B, nb], following
        GenMod.genphot (genmod.py:97-142) + ``sed``."""
        pp = np.asarray(photpars, dtype=np.float64)
        logt = np.log10(pp[:, 0])
        logl = 2.0 * pp[:, 4] + 4.0 * (logt - np.log10(5770.0))
        X = np.stack([10.0 ** logt, pp[:, 1], pp[:, 2], pp[:, 3], pp[:, 6], np.full(len(pp), 3.1)], 1)
        bc = self.bolcorr_batch(X)
        return (-2.5 * logl + 4.74 + 5.0 * np.log10(pp[:, 5]) - 5.0)[:, None] - bc

    def sed(self, logt=None, logg=None, feh=None, afe=None, logl=0.0, av=0.0,
            rv=3.1, dist=10.0, **extra):
        bc = self.bolcorr([10.0 ** logt, logg, feh, afe, av, rv])
        return -2.5 * logl + 4.74 - bc + 5.0 * np.log10(dist) - 5.0


class PayneSpecPredict(object):
    """Spectral MLP + broadening chain (YST-type network)."""

    def __init__(self, nnpath=None, Cnnpath=None):
        net = _load(nnpath)
        f8 = np.float64
        self.w = [net['w0'].astype(f8), net['w1'].astype(f8), net['w2'].astype(f8)]
        self.b = [net['b0'].astype(f8), net['b1'].astype(f8), net['b2'].astype(f8)]
        self.xmin = net['xmin'].astype(f8)
        self.xmax = net['xmax'].astype(f8)
        self.wavelength = net['wavelength'].astype(f8)
        self.resolution = float(net['resolution'])
        self.leak = float(net['leak'])
        self.logt_input = bool(net['logt_input'])
        self.d_in = self.w[0].shape[1]

    def _act(self, z):
        return np.where(z > 0, z, self.leak * z)

    def predictspec(self, Teff, logg, feh, afe, vmic=np.nan):
        lab = [np.log10(Teff) if self.logt_input else Teff, logg, feh, afe, vmic][:self.d_in]
        xs = (np.asarray(lab, dtype=np.float64) - self.xmin) / (self.xmax - self.xmin) - 0.5
        h = self._act(self.w[0] @ xs + self.b[0])
        h = self._act(self.w[1] @ h + self.b[1])
        return self.w[2] @ h + self.b[2]

    def getspec(self, Teff=None, logg=None, feh=None, afe=None, rad_vel=0.0,
                rot_vel=0.0, vmic=np.nan, inst_R=0.0, outwave=None, **extra):
        wave = self.wavelength
        flux = self.predictspec(Teff, logg, feh, afe, vmic)
        if rot_vel != 0.0:
            flux = sm.smooth_vsini(wave, flux, rot_vel, outwave=None, inres=0.0)
        if rad_vel != 0.0:
            wave = wave * (1.0 + rad_vel / C_KMS)
        if np.ndim(inst_R) == 0 and inst_R != 0.0:
            ow = None if outwave is None else np.asarray(outwave, dtype=np.float64)
            flux = sm.smooth_R(wave, flux, inst_R, ow, inres=self.resolution)
        elif outwave is not None:
            flux = np.interp(outwave, wave, flux, left=np.nan, right=np.nan)
        return wave, flux


def quantile(x, q, weights=None):
    """Weighted quantile by linear interpolation of the cumulative weights."""
    x = np.atleast_1d(x)
    q = np.atleast_1d(q)
    if weights is None:
        return np.percentile(x, list(100.0 * q))
    weights = np.atleast_1d(weights)
    idx = np.argsort(x)
    sw = weights[idx]
    cdf = np.cumsum(sw)[:-1]
    cdf /= cdf[-1]
    cdf = np.append(0, cdf)
    return np.interp(q, cdf, x[idx]).tolist()
