"""Host-side mirror of the reference's ``GenMod`` adaptor (reference
minesweeper/genmod.py:9-142): same method names and positional parameter lists,
evaluated by the CUDA kernels.  Scalar calls go through the batched entry
points with a batch of one; ``genphot_batch`` / ``genspec_batch`` take arrays.
"""
import numpy as np


class GenMod(object):
    def __init__(self, *arg, **kwargs):
        self.verbose = kwargs.get('verbose', False)
        self.engine = kwargs.get('engine', None)
        self.slot = kwargs.get('slot', 0)

    # the nets are loaded when the engine is built (likelihood.__init__); these
    # keep the reference's initialisation calls valid (genmod.py:15-44)
    def _initspecnn(self, nnpath=None, **kwargs):
        if self.engine is None or not hasattr(self.engine, 'spec_shape'):
            raise RuntimeError('engine has no spectral net loaded')

    def _initphotnn(self, filterarray=None, nnpath=None, **kwargs):
        if self.engine is None or not self.engine.bands:
            raise RuntimeError('engine has no photometric net loaded')
        self.filterarray = list(self.engine.bands)

    def genspec_batch(self, pars, modpoly=False):
        """pars[B, 8 + n_pc] -> flux[B, n_obs] (NumPy) on the star's obs_wave."""
        pars = np.atleast_2d(np.asarray(pars, dtype=np.float64))
        if not modpoly:
            pars = pars[:, :8]
        return self.engine.spec_model(pars, slot=self.slot).cpu().numpy()

    def genspec(self, pars, outwave=None, verbose=False, modpoly=False):
        """(modwave, modflux) like genmod.py:58-95; ``outwave`` must be the wavelength
        array registered for the star (the kernels resample onto it)."""
        flux = self.genspec_batch([[float(p) for p in pars]], modpoly=modpoly)[0]
        return outwave, flux

    def genphot_batch(self, pars):
        """pars[B,7] = (Teff, logg, feh, afe, logR, Dist, Av) -> mags[B, nb] (NumPy)."""
        return self.engine.phot_sed(np.atleast_2d(np.asarray(pars, dtype=np.float64))).cpu().numpy()

    def genphot(self, pars, rvfree=False, verbose=False):
        if rvfree:
            raise NotImplementedError('Rv is fixed at 3.1 on the batched path (genmod.py:106-109)')
        mags = self.genphot_batch([[float(p) for p in pars[:7]]])[0]
        return {ff: m for m, ff in zip(mags, self.engine.bands)}

    def genphot_scaled(self, pars, rvfree=False, verbose=False):
        raise NotImplementedError('photscale (log(A)) mode lives in the un-vendored Payne package '
                                  '(SURVEY.md Appendix B) and is out of scope')
