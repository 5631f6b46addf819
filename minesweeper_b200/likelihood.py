"""Drop-in mirror of the reference's ``likelihood`` class (reference
minesweeper/likelihood.py:7-219) whose arithmetic runs in the CUDA kernels.

Constructor signature, attributes (``fitpars_i``, ``ndim``, ``parsdict``, ``GM``,
``GMIST``, the ``*_bool`` flags, ``ageweight``, ``fixedpars``, ``dif_bool``) and the
methods ``lnlikefn(pars) -> float``, ``lnlike(specpars, photpars)`` and
``predMIST(dict)`` keep the reference's meaning, so ``fitstar.lnprobfn`` and the
prior object work unchanged.  The batched entry points a vectorised sampler uses
are ``lnlike_batch`` (arrays, stays on the GPU) and ``lnlikefn_batch`` (NumPy in /
out, plus one ``parsdict`` per row).

``fitargs['MISTpath' | 'photANNpath' | 'specANNpath']`` are model containers (dict
or ``.npz``); extra keyword ``precision`` in {'f64', 'f32', 'tc', 'tc16'} selects the
arithmetic mode and ``device`` the GPU.
"""
import numpy as np

from . import _lib
from .engine import Engine, load_container
from .fastMISTmod import GenMIST
from .genmod import GenMod

SPEC_KEYS = ('Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R')
FAIL_KEYS = ('Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'log(R)', 'log(Age)', 'Agewgt', 'Mass', 'log(L)')


def parsdict_keys(fitpars_i, fixedpars, iso=True, ageweight=True):
    """Key order of the ``parsdict`` the reference's ``lnlikefn`` leaves behind after a successful call
    (likelihood.py:95-141): sampled parameters, fixed parameters, then -- isochrone fits -- Teff, log(g), log(R),
    [Fe/H], [a/Fe], log(Age), Mass, log(L) and Agewgt, each appended only if not already present.  This is the
    column order of the samples file (fitstar.py:392-402)."""
    keys = list(fitpars_i)
    for k in fixedpars.keys():
        if k not in keys:
            keys.append(k)
    if iso:
        for k in ('Teff', 'log(g)', 'log(R)', '[Fe/H]', '[a/Fe]', 'EEP', 'log(Age)', 'Mass', 'log(L)'):
            if k not in keys:
                keys.append(k)
        if ageweight and 'Agewgt' not in keys:
            keys.append('Agewgt')
    return keys


class likelihood(object):
    def __init__(self, fitargs, fitpars, runbools, **kwargs):
        self.verbose = kwargs.get('verbose', True)
        self.fitargs = fitargs
        self.spec_bool, self.phot_bool, self.modpoly_bool, self.photscale_bool, self.flux_bool = \
            [bool(b) for b in runbools[:5]]
        if self.photscale_bool:
            raise NotImplementedError('photscale (log(A)) fits are not supported on the batched path')
        if self.flux_bool:
            raise NotImplementedError('fluxed-spectrum (contANNpath) fits are not supported')
        self.ageweight = self.fitargs['ageweight'] if 'ageweight' in fitargs else False
        self.fixedpars = self.fitargs['fixedpars']
        self.dif_bool = self.fitargs.get('mistdif', True)

        self.fitpars_i = [pp for pp in fitpars[0] if fitpars[1][pp]]
        self.ndim = len(self.fitpars_i)
        self.iso_bool = bool(fitpars[1].get('EEP', False))

        bands = list(fitargs['obs_phot'].keys()) if self.phot_bool else None
        self.engine = kwargs.get('engine', None)
        if self.engine is None:
            self.engine = Engine(
                mist=load_container(fitargs['MISTpath'], 'mist') if self.iso_bool else None,
                photnet=load_container(fitargs['photANNpath'], 'phot', bands=bands) if self.phot_bool else None,
                specnet=load_container(fitargs['specANNpath'], 'spec') if self.spec_bool else None,
                bands=bands, precision=kwargs.get('precision', 'f32'),
                device=kwargs.get('device', None))
        self.slot = kwargs.get('slot', 0)
        self.engine.set_star(
            slot=self.slot,
            obs_phot=fitargs['obs_phot'] if self.phot_bool else None,
            obs_wave=fitargs['obs_wave_fit'] if self.spec_bool else None,
            obs_flux=fitargs['obs_flux_fit'] if self.spec_bool else None,
            obs_eflux=fitargs['obs_eflux_fit'] if self.spec_bool else None)

        # array-valued Inst_R (fixed): the wavelength-dependent LSF mode of genspec (genmod.py:74-77)
        self.lsf_sigma = None
        inst = self.fixedpars.get('Inst_R', None) if isinstance(self.fixedpars, dict) else None
        if self.spec_bool and inst is not None and np.ndim(inst) > 0:
            self.lsf_sigma = np.asarray(inst, dtype=np.float64)
            self.engine.set_lsf(self.slot, self.lsf_sigma)
            self.fixedpars = dict(self.fixedpars)
            self.fixedpars['Inst_R'] = np.nan              # the scalar slot of the parameter block is unused in LSF mode
        self.GM = GenMod(engine=self.engine, slot=self.slot)
        if self.spec_bool:
            self.GM._initspecnn(nnpath=fitargs.get('specANNpath'), NNtype=fitargs.get('NNtype'))
        if self.phot_bool:
            self.GM._initphotnn(bands, nnpath=fitargs.get('photANNpath'))
        if self.iso_bool:
            self.GMIST = GenMIST(MISTpath=fitargs['MISTpath'], ageweight=self.ageweight,
                                 engine=self.engine, verbose=False)

        self.pc_names = [pp for pp in self.fitpars_i if 'pc' in pp]
        self._colmap = Engine.make_colmap(self.fitpars_i)
        self._fixed = Engine.make_fixed(self.fixedpars)
        self._flags = Engine.make_flags(spec=self.spec_bool, phot=self.phot_bool,
                                        modpoly=self.modpoly_bool, ageweight=self.ageweight,
                                        mistdif=self.dif_bool, n_pc=len(self.pc_names), slot=self.slot)
        self.parsdict = {}
        self.MISTrename = {'log_age': 'log(Age)', 'star_mass': 'Mass', 'log_R': 'log(R)',
                           'log_L': 'log(L)', 'log_Teff': 'log(Teff)', 'log_g': 'log(g)'}

    # ---- batched entry points ---------------------------------------------------
    def lnlike_batch(self, theta, want_derived=True, want_brackets=False, star_slot=None):
        """theta[B, ndim] in ``fitpars_i`` column order (GPU tensor or array).
        Returns GPU tensors (lnL[B], derived[B,13] | None, brackets[B,4] | None)."""
        return self.engine.lnlike_batch(theta, self._colmap, self._fixed, self._flags,
                                        star_slot=star_slot, want_derived=want_derived,
                                        want_brackets=want_brackets)

    def lnlike_batch_host(self, theta, want_derived=False, **kw):
        """NumPy in, NumPy out through the host-buffer C-ABI call (copies included)."""
        return self.engine.lnlike_batch_host(theta, self._colmap, self._fixed, self._flags,
                                             want_derived=want_derived, **kw)

    def parsdict_from(self, pars, derived_row):
        """The dict ``lnlikefn`` leaves in ``self.parsdict`` (likelihood.py:95-141)."""
        pd = {pp: vv for pp, vv in zip(self.fitpars_i, pars)}
        for kk in self.fixedpars.keys():
            pd[kk] = self.fixedpars[kk]
        if derived_row is not None:
            d = dict(zip(_lib.DERIVED_NAMES, derived_row))
            if not np.isfinite(d['log(Teff)']):
                for kk in FAIL_KEYS:
                    pd[kk] = np.inf
                return pd
            pd['Teff'] = 10.0 ** d['log(Teff)']
            pd['log(g)'] = d['log(g)']
            pd['log(R)'] = d['log(R)']
            if self.dif_bool:
                pd['[Fe/H]'], pd['[a/Fe]'] = d['[Fe/H]'], d['[a/Fe]']
            else:
                pd['[Fe/H]'], pd['[a/Fe]'] = d['initial_[Fe/H]'], d['initial_[a/Fe]']
            pd['EEP'] = d['EEP']
            pd['log(Age)'] = d['log(Age)']
            pd['Mass'] = d['Mass']
            pd['log(L)'] = d['log(L)']
            if self.ageweight:
                pd['Agewgt'] = d['Agewgt']
        return pd

    def lnlikefn_batch(self, theta):
        """theta[B, ndim] -> (lnL[B] NumPy, [parsdict_0, ...])."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        lnl, der, _ = self.lnlike_batch(theta, want_derived=self.iso_bool)
        lnl = lnl.cpu().numpy()
        der = der.cpu().numpy() if der is not None else [None] * len(theta)
        return lnl, [self.parsdict_from(t, d) for t, d in zip(theta, der)]

    # ---- the reference's scalar API ---------------------------------------------
    def lnlikefn(self, pars):
        lnl, dicts = self.lnlikefn_batch([[float(p) for p in pars]])
        self.parsdict = dicts[0]
        return float(lnl[0])

    def predMIST(self, inpars):
        pred = self.GMIST.getMIST(eep=inpars['EEP'], mass=inpars['initial_Mass'],
                                  feh=inpars['initial_[Fe/H]'], afe=inpars['initial_[a/Fe]'],
                                  verbose=False)
        if pred is None:
            return None
        return {kk: pp for kk, pp in zip(self.GMIST.modpararr, pred)}

    def lnlike(self, specpars=None, photpars=None):
        """-0.5 (chi2_spec + chi2_phot) for explicit parameter lists (likelihood.py:183-219)."""
        chi = 0.0
        if self.spec_bool:
            _, mod = self.GM.genspec(specpars, outwave=self.fitargs['obs_wave_fit'],
                                     modpoly=self.modpoly_bool)
            o, s = self.fitargs['obs_flux_fit'], self.fitargs['obs_eflux_fit']
            chi += np.nansum((mod - o) ** 2.0 / s ** 2.0)
        if self.phot_bool:
            sed = self.GM.genphot(photpars)
            op = self.fitargs['obs_phot']
            chi += np.nansum([(sed[kk] - op[kk][0]) ** 2.0 / op[kk][1] ** 2.0 for kk in op.keys()])
        return -0.5 * chi
