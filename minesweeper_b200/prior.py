"""Host-side mirror of the reference's ``prior`` class (reference minesweeper/prior.py:5-639)
for the batched path: same constructor ``prior(fitargs, inpriordict, fitpars, runbools)``, same
``priortrans(upars)`` / ``lnpriorfn(pars)`` scalar API, plus ``priortrans_batch`` and
``lnprob_batch`` that run as CUDA kernels (csrc/prior.cu) on whole batches.

The reference's precedence rules are reproduced when the priordict is translated:
volume priors (``pv_*``) choose the transform of a sampled parameter in the order
uniform > gaussian > tgaussian > [beta] > exp > ... > default range (prior.py:176-385);
every other entry is an additive term applied per parameter family (prior.py:519-639).
``pv_beta`` transforms, the Galactic distance transform (``GAL`` / ``GALAGE``: ``1000 AP.gal_ppf(u)``,
prior.py:333-336) and the advanced additive priors ``GAL``, ``GALAGE``, ``VROT``, ``ALPHA`` (prior.py:413-501 ->
advancedpriors.py:410-833) run in the same kernels.  ``VTOT`` and ``AngDia`` are accepted and have no effect, exactly
as in the reference: ``VTOT`` sets ``pm_bool`` while ``lnpriorfn`` tests ``vtot_bool`` (prior.py:26, 84, 460), and
``AngDia`` only reaches the ``AdvancedPriors`` constructor (prior.py:30-33, 144-145) -- ``AngDia_lnprior`` is never
called.  Unsupported (constructor raises): ``pv_texp`` (NameError in the reference itself, prior.py:367) and the IMF
prior without ``initial_Mass`` (precedence bug at prior.py:404).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import PriorCol, PriorTerm, check

MIST_PARS = ['EEP', 'initial_Mass', 'initial_[Fe/H]', 'initial_[a/Fe]']
SPEC_PARS = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R']
ISO_PHOT_PARS = ['log(A)', 'log(R)', 'Av', 'Rv', 'Dist']
DEFAULTS = {                                                         # prior.py:115-135
    'Teff': [3000.0, 17000.0], 'log(g)': [-1.0, 5.5], '[Fe/H]': [-4.0, 0.5], '[a/Fe]': [-0.2, 0.6],
    'Vrad': [-700.0, 700.0], 'Vrot': [0, 30.0], 'Vmic': [0.5, 3.0], 'Inst_R': [10000.0, 60000.0],
    'log(A)': [-3.0, 7.0], 'log(R)': [-2.0, 3.0], 'Dist': [1.0, 200000.0], 'Av': [0.0, 5.0],
    'Rv': [2.0, 5.0], 'EEP': [200, 808], 'initial_Mass': [0.25, 30.0], 'initial_[Fe/H]': [-4.0, 0.5],
    'initial_[a/Fe]': [-0.2, 0.6], 'log(Age)': [9.0, 10.2], 'Age': [1.0, 14.0]}
NO_EFFECT = ('VTOT', 'AngDia')          # accepted, never applied by the reference's lnpriorfn (see module docstring)
VALUE_IDS = {'log(Age)': _lib.MS_V_LOGAGE, 'Age': _lib.MS_V_AGE, 'Parallax': _lib.MS_V_PARALLAX}


class prior(object):
    def __init__(self, fitargs, inpriordict, fitpars, runbools, engine=None, likeobj=None):
        self.fitargs = fitargs
        self.fixedpars = fitargs['fixedpars']
        self.fitpars_i = [pp for pp in fitpars[0] if fitpars[1][pp]]
        self.ndim = len(self.fitpars_i)
        self.spec_bool, self.phot_bool, self.modpoly_bool, self.photscale_bool = [bool(b) for b in runbools[:4]]
        self.engine = engine if engine is not None else (likeobj.engine if likeobj is not None else None)
        if self.engine is None:
            raise ValueError('prior needs the Engine of the likelihood object (engine= or likeobj=)')
        self.imf_bool = self.gal_bool = self.galage_bool = self.vrot_bool = self.alpha_bool = False
        self.priordict = {k: {} for k in ('uniform', 'gaussian', 'tgaussian', 'exp', 'texp', 'loguniform', 'beta')}
        self.additionalpriors = {}
        self.polycoefarr = None
        for kk, spec in inpriordict.items():
            if kk == 'blaze_coeff':
                self.polycoefarr = spec
            elif kk == 'IMF':
                self.imf_bool = True
                self.imf = spec['IMF_type']
            elif kk in ('GAL', 'GALAGE'):                               # prior.py:56-75
                self.gal_bool = True
                self.galage_bool = self.galage_bool or kk == 'GALAGE'
                self.lb_coords = spec['lb_coords']
                if kk == 'GALAGE':
                    self.agepars = spec['pars']
                dd = inpriordict.get('Dist', {})
                self.mindist, self.maxdist = (dd['pv_uniform'][0], dd['pv_uniform'][1]) if 'Dist' in inpriordict \
                    else (1.0, 200000.0)
            elif kk == 'VROT':
                self.vrot_bool, self.vrotpars = True, spec
            elif kk == 'ALPHA':
                self.alpha_bool, self.minalpha = True, spec['min']
            elif kk in NO_EFFECT:
                pass
            else:
                for ii, val in spec.items():
                    if ii.startswith('pv_'):
                        self.priordict[ii[3:]][kk] = val
                    elif ii != 'fixed':
                        self.additionalpriors.setdefault(kk, {})[ii] = val
        if self.priordict['texp']:
            raise NotImplementedError('pv_texp raises NameError in the reference itself (prior.py:367): not offered')
        self.iso_bool = 'EEP' in self.fitpars_i or 'EEP' in self.fixedpars
        if self.imf_bool and 'initial_Mass' not in self.fitpars_i and 'initial_Mass' not in self.fixedpars:
            raise NotImplementedError('IMF prior without initial_Mass (prior.py:404 path) is not supported')
        self._adv = None
        if self.gal_bool or self.vrot_bool or self.alpha_bool:
            from . import advancedpriors as AP
            adv = _lib.AdvPrior()
            adv.gal, adv.galage = int(self.gal_bool), int(self.galage_bool)
            adv.vrot, adv.alpha = int(self.vrot_bool), int(self.alpha_bool)
            if self.gal_bool:
                adv.Xp, adv.Yp, adv.Zp = AP.sight_line(self.lb_coords[0], self.lb_coords[1])
                adv.sol_X, adv.sol_Z = AP.SOL_X, AP.SOL_Z
                if 'Dist' in self.fitpars_i and hasattr(self.engine, 'lib'):   # distance transform table (prior.py:333-336)
                    x, cdf = AP.distance_table(self.lb_coords[0], self.lb_coords[1], self.mindist / 1000.0,
                                               self.maxdist / 1000.0)
                    check(self.engine.lib.ms_set_prior_table(self.engine._ctx, x.ctypes.data_as(_lib.c_dp),
                                                             cdf.ctypes.data_as(_lib.c_dp), len(x)))
            if self.galage_bool:                                        # advancedpriors.py:795-797 defaults
                defaults = dict(thin={'min': 4.0, 'max': 14.0}, thick={'min': 6.0, 'max': 14.0, 'mean': 10.0, 'sigma': 2.0},
                                halo={'min': 8.0, 'max': 14.0, 'mean': 12.0, 'sigma': 2.0})
                for comp, field in (('thin', adv.age_thin), ('thick', adv.age_thick), ('halo', adv.age_halo)):
                    pp = self.agepars.get(comp, defaults[comp])
                    field[0], field[1] = float(pp['min']), float(pp['max'])
                    field[2], field[3] = (float(pp['mean']), float(pp['sigma'])) if 'mean' in pp else (0.0, -1.0)
            if self.vrot_bool:
                for key, field in (('dwarf', adv.vrot_dwarf), ('giant', adv.vrot_giant)):
                    field[0], field[1], field[2] = (float(self.vrotpars[key][k]) for k in ('a', 'c', 'n'))
                if 'Vrot' not in self.fitpars_i and 'Vrot' not in self.fixedpars:
                    raise KeyError("VROT prior needs 'Vrot' (the reference raises KeyError)")
            if self.alpha_bool:
                adv.minalpha = float(self.minalpha)
            self._adv = adv
        self._cols = (PriorCol * self.ndim)(*[self._column(n) for n in self.fitpars_i])
        terms = self._terms()
        if len(terms) > _lib.MS_MAX_PRIOR_TERMS:
            raise ValueError('too many additive prior terms')
        self._nterms = len(terms)
        self._terms_c = (PriorTerm * max(len(terms), 1))(*terms)
        self._colmap = np.array([_lib.param_id(p) for p in self.fitpars_i], dtype=np.int32)
        self._fixed = np.full(_lib.MS_NPARAM, np.nan)
        for k, v in self.fixedpars.items():
            try:
                self._fixed[_lib.param_id(k)] = float(v)
            except KeyError:
                pass
        self._flags = _lib.Flags(int(self.spec_bool), int(self.phot_bool), int(self.modpoly_bool),
                                 int(fitargs.get('ageweight', False)), int(fitargs.get('mistdif', True)), 0)

    # ---- translation of the reference priordict ------------------------------------------
    def _column(self, name):
        pd = self.priordict
        if name.startswith('pc_'):                                   # prior.py:262-288
            loc, scale = self.polycoefarr[int(name.split('_')[-1])][:2]
            return _col('tgaussian', loc - 5.0 * scale, loc + 5.0 * scale, loc, scale)
        if name == 'Dist' and self.gal_bool and self.phot_bool:        # prior.py:331-336: 1000 AP.gal_ppf(u)
            return _col('table', 1000.0)
        if name in MIST_PARS:                                        # prior.py:176-209
            order = ('uniform', 'gaussian', 'tgaussian', 'beta')
        elif name in SPEC_PARS:                                      # :212-246, and :297-327 for SED-only fits
            order = ('uniform', 'gaussian', 'tgaussian', 'beta', 'exp')
        elif name in ISO_PHOT_PARS:                                  # :338-381 (texp not offered)
            order = ('uniform', 'gaussian', 'exp', 'tgaussian', 'beta', 'loguniform')
        else:
            raise KeyError('no prior transform rule for %r' % name)
        for kind in order:
            if name in pd[kind]:
                v = pd[kind][name]
                if kind == 'uniform':
                    return _col('uniform', min(v), max(v))
                return _col(kind, *v)
        return _col('uniform', DEFAULTS[name][0], DEFAULTS[name][1])

    def _terms(self):
        out = []
        for kk, spec in self.additionalpriors.items():
            fams = []
            if self.iso_bool and kk in MIST_PARS + ['log(Age)', 'Age']:
                fams.append('mist')
            if self.spec_bool and kk in SPEC_PARS:
                fams.append('spec')
            if self.phot_bool:
                phot_keys = ([] if self.spec_bool else ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]'])
                phot_keys += [k for k in ('log(R)', 'Dist', 'log(A)', 'Av') if k in self.fitpars_i]
                if 'Dist' in self.fitpars_i:
                    phot_keys.append('Parallax')
                if kk in phot_keys:
                    fams.append('phot')
            if not fams:
                raise KeyError('additive prior on %r does not apply to this fit (the reference raises KeyError)' % kk)
            vid = VALUE_IDS.get(kk)
            if vid is None:
                vid = _lib.param_id(kk)
            for fam in fams:
                for ii, v in spec.items():
                    if ii == 'uniform':
                        out.append(_term(vid, 'uniform', v[0], v[1]))
                    elif ii == 'gaussian':
                        out.append(_term(vid, 'gaussian', v[0], v[1]))
                    elif ii == 'tgaussian':
                        out.append(_term(vid, 'tgaussian' if fam == 'mist' else 'tgaussian_bounds', *v))
                    else:
                        raise NotImplementedError('%s prior on %s' % (ii, kk))
        return out

    # ---- batched entry points --------------------------------------------------------------
    def priortrans_batch(self, u, out=None):
        """u[B, ndim] in the unit cube (GPU tensor or array) -> theta[B, ndim] GPU tensor."""
        eng = self.engine
        ut = eng._dev(u)
        B = ut.shape[0]
        theta = out if out is not None else torch.empty_like(ut)
        with torch.cuda.device(eng.device):
            check(eng.lib.ms_prior_transform(eng._ctx, ut.data_ptr(), B, self.ndim, self._cols,
                                             theta.data_ptr(), eng._stream()))
        return theta

    def lnprob_batch(self, theta, lnl=None, derived=None, out=None, star_slot=None, star_plx=None):
        """lnL + lnprior with lnprobfn's -inf short circuits (fitstar.py:715-727); with
        ``lnl=None`` the ln-prior alone.  ``derived`` is the [B,13] block of lnlike_batch.
        Catalog mode: ``star_slot[B]`` (int32 GPU tensor) and ``star_plx[nslots,2]`` (f64 GPU tensor)
        add a per-star Gaussian parallax prior."""
        eng = self.engine
        th = eng._dev(theta)
        B = th.shape[0]
        if out is None:
            out = torch.empty((B,), dtype=torch.float64, device=eng.device)
        if self.iso_bool and derived is None:
            raise ValueError('isochrone fits need the derived block for the ln-prior')
        with torch.cuda.device(eng.device):
            check(eng.lib.ms_lnprob_batch_adv(
                eng._ctx, C.byref(self._flags), th.data_ptr(), B, self.ndim,
                self._colmap.ctypes.data_as(_lib.c_ip), self._fixed.ctypes.data_as(_lib.c_dp),
                derived.data_ptr() if derived is not None else None, self._terms_c, self._nterms,
                int(self.imf_bool), lnl.data_ptr() if lnl is not None else None, out.data_ptr(),
                star_slot.data_ptr() if star_slot is not None else None,
                star_plx.data_ptr() if star_plx is not None else None,
                C.byref(self._adv) if self._adv is not None else None, eng._stream()))
        return out

    # ---- the reference's scalar API ---------------------------------------------------------
    def priortrans(self, upars):
        th = self.priortrans_batch(np.asarray([[float(x) for x in upars]]))
        return [float(v) for v in th.cpu().numpy()[0]]

    def lnpriorfn(self, pars):
        pd = dict(zip(self.fitpars_i, pars)) if isinstance(pars, list) else dict(pars)
        pd.update(self.fixedpars)
        theta = np.array([[float(pd[k]) for k in self.fitpars_i]])
        der = None
        if self.iso_bool:
            g = lambda k: float(pd.get(k, np.nan))
            teff = g('Teff')
            der = np.array([[g('EEP'), g('initial_Mass'), g('[Fe/H]'), g('[a/Fe]'), g('log(Age)'), g('Mass'),
                             g('log(R)'), g('log(L)'), np.log10(teff) if teff > 0 else np.nan, g('[Fe/H]'),
                             g('[a/Fe]'), g('log(g)'), g('Agewgt')]])
            der = self.engine._dev(der)
        return float(self.lnprob_batch(theta, None, der).cpu().numpy()[0])


def _col(kind, *p):
    c = PriorCol()
    c.kind = _lib.PRIOR_KINDS[kind]
    for i, v in enumerate(p[:4]):
        c.p[i] = float(v)
    return c


def _term(vid, kind, *p):
    t = PriorTerm()
    t.value_id = int(vid)
    t.kind = _lib.TERM_KINDS[kind]
    for i, v in enumerate(p[:4]):
        t.p[i] = float(v)
    return t
