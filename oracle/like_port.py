"""ORACLE (test infrastructure only) -- per-point NumPy f64 restatement of the
reference's likelihood object and its forward-model adaptor.

Follows
  reference minesweeper/likelihood.py:9-69   constructor contract
                                     :71-90   predMIST
                                     :92-181  lnlikefn (parsdict side effect, -inf / +inf conventions)
                                     :183-219 lnlike (nansum chi-square over pixels and observed bands)
  reference minesweeper/genmod.py:58-95      genspec (Inst_R * 2.355, blaze polynomial)
                                  :97-142     genphot (logL from log R and Teff, Rv = 3.1)
  reference minesweeper/fitutils.py:11-20    polycalc (Chebyshev on wavelength rescaled to [-1, 1])
  reference minesweeper/fitstar.py:715-727   lnprobfn
Pinned against the reference classes themselves (run through oracle/refshim.py
with oracle/payne_port.py standing in for the un-vendored Payne package) by
tests/golden/like_*.npz.
"""
import numpy as np
from numpy.polynomial.chebyshev import chebval

from .mist_port import GenMISTPort, dense_to_rows
from .payne_port import FastPayneSEDPredict, PayneSpecPredict, _load

SPEC_KEYS = ('Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R')
FAIL_KEYS = ('Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'log(R)', 'log(Age)', 'Agewgt', 'Mass', 'log(L)')


def blaze(coef, wave):
    """Chebyshev series on this spectrum's own min/max (fitutils.py:11-20)."""
    x = wave - wave.min()
    x = 2.0 * (x / x.max()) - 1.0
    return chebval(x, coef)


def _mist_rows(container):
    c = _load(container)
    if 'values' in c:                       # dense form -> track tables
        return dense_to_rows(c)
    return c


class LikelihoodPort(object):
    def __init__(self, fitargs, fitpars, runbools, **kw):
        self.fitargs = fitargs
        (self.spec_bool, self.phot_bool, self.modpoly_bool,
         self.photscale_bool, self.flux_bool) = runbools[:5]
        self.ageweight = fitargs['ageweight']
        self.fixedpars = fitargs['fixedpars']
        self.dif_bool = fitargs['mistdif']
        if self.spec_bool:
            self.PP = PayneSpecPredict(nnpath=fitargs['specANNpath'])
        if self.phot_bool:
            self.fppsed = FastPayneSEDPredict(usebands=fitargs['obs_phot'].keys(),
                                              nnpath=fitargs['photANNpath'])
            self.filterarray = self.fppsed.filternames
        if fitpars[1]['EEP']:
            self.GMIST = GenMISTPort(_mist_rows(fitargs['MISTpath']), ageweight=self.ageweight)
        self.fitpars_i = [p for p in fitpars[0] if fitpars[1][p]]
        self.ndim = len(self.fitpars_i)
        self.parsdict = {}

    # -- likelihood.py:92-181 -------------------------------------------
    def lnlikefn(self, pars):
        pd = dict(zip(self.fitpars_i, pars))
        pd.update(self.fixedpars)
        self.parsdict = pd
        if 'EEP' in pd:
            m = self.GMIST.getMIST(eep=pd['EEP'], mass=pd['initial_Mass'],
                                   feh=pd['initial_[Fe/H]'], afe=pd['initial_[a/Fe]'])
            if m is None:
                for k in FAIL_KEYS:
                    pd[k] = np.inf
                return -np.inf
            md = dict(zip(self.GMIST.modpararr, m))
            pd['Teff'] = 10.0 ** md['log(Teff)']
            pd['log(g)'] = md['log(g)']
            pd['log(R)'] = md['log(R)']
            if self.dif_bool:
                pd['[Fe/H]'], pd['[a/Fe]'] = md['[Fe/H]'], md['[a/Fe]']
            else:
                pd['[Fe/H]'], pd['[a/Fe]'] = md['initial_[Fe/H]'], md['initial_[a/Fe]']
            pd['EEP'] = md['EEP']
            pd['log(Age)'] = md['log(Age)']
            pd['Mass'] = md['Mass']
            pd['log(L)'] = md['log(L)']
            if self.ageweight:
                pd['Agewgt'] = md['Agewgt']
        specpars = photpars = None
        if self.spec_bool:
            specpars = [pd.get(k, np.nan) for k in SPEC_KEYS]
            if self.modpoly_bool:
                specpars += [pd[k] for k in self.fitpars_i if 'pc' in k]
        if self.phot_bool:
            photpars = [pd[k] for k in ('Teff', 'log(g)', '[Fe/H]', '[a/Fe]')]
            if 'log(A)' in self.fitpars_i:
                photpars += [pd['log(A)']]
            else:
                photpars += [pd['log(R)'], pd['Dist']]
            photpars += [pd['Av']]
        lnl = self.lnlike(specpars=specpars, photpars=photpars)
        if self.ageweight:
            with np.errstate(invalid='ignore', divide='ignore'):
                lnl = lnl + np.log(pd['Agewgt'])
        return lnl

    # -- genmod.py:58-95 ------------------------------------------------
    def genspec(self, pars, outwave=None, modpoly=False):
        teff, logg, feh, afe, rv, vrot, vmic, inst_r = pars[:8]
        inst = 2.355 * inst_r if isinstance(inst_r, float) else inst_r
        wave, flux = self.PP.getspec(Teff=teff, logg=logg, feh=feh, afe=afe,
                                     rad_vel=rv, rot_vel=vrot, vmic=vmic,
                                     inst_R=inst, outwave=outwave)
        if outwave is None:
            outwave = wave
        if modpoly:
            flux = flux * blaze(pars[8:], outwave)
        return wave, flux

    # -- genmod.py:97-142 -----------------------------------------------
    def genphot(self, pars):
        teff, logg, feh, afe, logr, dist, av = pars[:7]
        logt = np.log10(teff)
        logl = 2.0 * logr + 4.0 * (logt - np.log10(5770.0))
        mags = self.fppsed.sed(logt=logt, logg=logg, feh=feh, afe=afe,
                               logl=logl, av=av, dist=dist, rv=3.1)
        return dict(zip(self.filterarray, mags))

    # -- likelihood.py:183-219 ------------------------------------------
    def lnlike(self, specpars=None, photpars=None):
        chi_spec = chi_phot = 0.0
        if self.spec_bool:
            _, mod = self.genspec(specpars, outwave=self.fitargs['obs_wave_fit'],
                                  modpoly=self.modpoly_bool)
            obs, err = self.fitargs['obs_flux_fit'], self.fitargs['obs_eflux_fit']
            chi_spec = np.nansum((mod - obs) ** 2.0 / err ** 2.0)
        if self.phot_bool:
            if self.photscale_bool:
                raise NotImplementedError('photscale (log(A)) mode: formula lives in the '
                                          'un-vendored Payne package (SURVEY.md Appendix B)')
            sed = self.genphot(photpars)
            op = self.fitargs['obs_phot']
            chi_phot = np.nansum([(sed[k] - op[k][0]) ** 2.0 / op[k][1] ** 2.0 for k in op.keys()])
        return -0.5 * (chi_spec + chi_phot)


def lnprobfn(pars, likeobj, priorobj):
    """fitstar.py:715-727."""
    lnl = likeobj.lnlikefn(pars)
    if lnl == -np.inf:
        return -np.inf
    lnp = priorobj.lnpriorfn(likeobj.parsdict)
    if lnp == -np.inf:
        return -np.inf
    return lnp + lnl


def lnlike_phot_dense(mist, photnet, obs_phot, theta6, ageweight=True, mistdif=True, rel=1e-5):
    """Batched restatement of the photometry-only isochrone likelihood (config 2) on the dense grid
    container: the vectorised ``mist_port.interp_dense`` (pinned to the reference's getMIST by the goldens)
    -> derived parameters (likelihood.py:121-141) -> ``FastPayneSEDPredict.sed_batch`` (genmod.py:97-142)
    -> nansum chi-square over the observed bands (likelihood.py:207-210) -> + log(Agewgt) (:168-176).
    theta6 columns = (Dist, Av, EEP, initial_[Fe/H], initial_[a/Fe], initial_Mass) = ``fitpars_i`` of that
    fit.  For parity checks at sizes where the per-point form is too slow (bench.py's in-run sample, the
    full-grid GPU test).  Returns a dict with brackets[B,4], ok[B], pred[B,9], mags[B,nb] (NaN rows when
    not ok), lnl[B] (-inf when not ok) and ``bound[B]``: the |d lnL| a relative model-magnitude error
    ``rel`` can cause (first + second order), the per-row tolerance of the fp32-class modes."""
    from .mist_port import interp_dense
    th = np.asarray(theta6, dtype=np.float64)
    dist, av, eep, feh, afe, mass = th.T
    br, pred, ok = interp_dense(mist, eep, mass, feh, afe)
    sed = FastPayneSEDPredict(usebands=list(obs_phot.keys()), nnpath=photnet)
    # columns of pred: log(Age), Mass, log(R), log(L), log(Teff), [Fe/H], [a/Fe], log(g), Agewgt
    teff = 10.0 ** pred[:, 4]
    sfeh, safe_ = (pred[:, 5], pred[:, 6]) if mistdif else (feh, afe)
    pp = np.stack([teff, pred[:, 7], sfeh, safe_, pred[:, 2], dist, av], 1)
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        mags = np.full((len(th), len(sed.filternames)), np.nan)
        mags[ok] = sed.sed_batch(pp[ok])
        o = np.array([obs_phot[b][0] for b in sed.filternames])
        s = np.array([obs_phot[b][1] for b in sed.filternames])
        chi = np.nansum((mags - o) ** 2.0 / s ** 2.0, axis=1)
        lnl = -0.5 * chi
        if ageweight:
            lnl = lnl + np.log(pred[:, 8])
        lnl = np.where(ok, lnl, -np.inf)
        dm = rel * np.abs(mags)
        bound = np.nansum((np.abs(mags - o) * dm + 0.5 * dm ** 2) / s ** 2, axis=1)
    return dict(brackets=br, ok=ok, pred=pred, mags=mags, lnl=lnl, bound=bound, photpars=pp,
                bands=list(sed.filternames), obs_mag=o, obs_err=s)
