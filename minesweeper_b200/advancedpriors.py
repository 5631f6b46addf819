"""Host-side pieces of the reference's ``AdvancedPriors`` (reference minesweeper/advancedpriors.py) that the batched
prior needs at construction time.  The per-point arithmetic of the Galactic / rotation / alpha priors runs in
``lnprior_kernel`` (csrc/prior.cu); what is left for the host is the one-time table behind the Galactic distance
transform:

    AdvancedPriors.__init__   distarr = logspace(log10(mindist), log10(maxdist), 10000),
                              distnormfactor = exp(gal_lnprior(distarr))              advancedpriors.py:43-63
    AdvancedPriors.gal_ppf    Payne's weighted quantile of distarr with those weights     advancedpriors.py:665-670

``sight_line`` gives the direction cosines the kernel uses (advancedpriors.py:46-56); ``gal_lnprior`` is the
``coords=[]`` branch of the reference's function (no astropy), vectorised, returning ln p(d) and the three number-density
components (thin disk, thick disk, halo).
"""
import numpy as np

SOL_X, SOL_Z = 8.3, -27.0 / 1000.0                                   # advancedpriors.py:55-56


def sight_line(l_deg, b_deg):
    lp, bp = np.deg2rad(l_deg), np.deg2rad(b_deg)
    return float(np.cos(lp) * np.cos(bp)), float(np.sin(lp) * np.cos(bp)), float(np.sin(bp))


def _logn_disk(R, Z, R_scale, Z_scale, R_solar=8.2, Z_solar=0.025):
    return -((R - R_solar) / R_scale + (np.abs(Z) - np.abs(Z_solar)) / Z_scale)


def _logn_halo(R, Z, R_solar=8.2, Z_solar=0.025, R_smooth=0.5, eta=4.2, q_ctr=0.2, q_inf=0.8, r_q=6.0):
    r = np.sqrt(R ** 2 + Z ** 2)
    q = q_inf - (q_inf - q_ctr) * np.exp(1.0 - np.sqrt(r ** 2 + r_q ** 2) / r_q)
    Reff = np.sqrt(R ** 2 + (Z / q) ** 2 + R_smooth ** 2)
    q_sol = q_inf - (q_inf - q_ctr) * np.exp(1.0 - np.sqrt(R_solar ** 2 + Z_solar ** 2 + r_q ** 2) / r_q)
    Reff_sol = np.sqrt(R_solar ** 2 + (Z_solar / q_sol) + R_smooth ** 2)      # un-squared term as in the reference (:322)
    return -eta * np.log(Reff / Reff_sol)


def gal_lnprior(dists, l_deg, b_deg):
    """ln prior of the distance(s) in kpc along (l, b) and the [thin, thick, halo] ln number densities
    (advancedpriors.py:410-558 with the default structural parameters, coords = [])."""
    d = np.asarray(dists, dtype=np.float64)
    Xp, Yp, Zp = sight_line(l_deg, b_deg)
    vol = 2.0 * np.log(d + 1e-300)
    X, Y, Z = d * Xp - SOL_X, d * Yp, d * Zp - SOL_Z
    R = np.hypot(X, Y)
    thin = _logn_disk(R, Z, 2.6, 0.3) + vol
    thick = _logn_disk(R, Z, 2.0, 0.9) + vol + np.log(0.04)
    halo = _logn_halo(R, Z) + vol + np.log(0.005)
    comp = np.stack([thin, thick, halo])
    m = comp.max(axis=0)
    return m + np.log(np.exp(comp - m).sum(axis=0)), comp


def distance_table(l_deg, b_deg, mindist_kpc, maxdist_kpc, n=10000):
    """(x[n], cdf[n]) such that ``np.interp(u, cdf, x)`` is ``AdvancedPriors.gal_ppf(u)`` (kpc)."""
    x = np.logspace(np.log10(mindist_kpc), np.log10(maxdist_kpc), n)
    w = np.exp(gal_lnprior(x, l_deg, b_deg)[0])
    cdf = np.cumsum(w)[:-1]
    cdf = cdf / cdf[-1]
    return x, np.append(0.0, cdf)
