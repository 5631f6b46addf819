"""Seeded synthetic model files of the real shape.

The reference's model files (MIST EEP tracks, Payne ANN weights) are not
shipped (reference data/MIST/README.md:1, data/nnMIST/README.md:1), so every
test and benchmark in this repo runs on the generators below.  Three kinds of
model are produced, each as a plain ``dict`` of NumPy arrays that can be saved
with ``np.savez`` and handed to ``fitargs['MISTpath'|'photANNpath'|'specANNpath']``:

* ``make_mist``      -- dense MIST grid + validity mask (missing / truncated tracks)
* ``make_phot_net``  -- per-band 6 -> H -> H -> 1 sigmoid MLP stack (Payne FastPayneSEDPredict shape)
* ``make_spec_net``  -- D -> H -> H -> P leaky-ReLU MLP on a log-uniform wavelength grid (Payne YST shape)

``mist_rows`` converts the dense grid into the HDF5-like row tables the
reference reads (reference fastMISTmod.py:76-98), which is what the oracle and
the reference shim consume.
"""
import numpy as np

# Output-column order of GenMIST.output (reference fastMISTmod.py:48-60, 84-87)
MIST_COLUMNS = ('log_age', 'star_mass', 'log_R', 'log_L', 'log_Teff',
                '[Fe/H]', '[a/Fe]', 'log_g', 'Agewgt')
# Field names of one HDF5 track table (reference fastMISTmod.py:16-23, 80)
MIST_LABEL_FIELDS = ('EEP', 'initial_mass', 'initial_[Fe/H]', 'initial_[a/Fe]')

# The ten finite bands of the demo star (reference demo/phot/18Sco_phot.dat)
DEMO_BANDS = ('2MASS_J', '2MASS_H', '2MASS_Ks', 'GaiaEDR3_G', 'GaiaEDR3_BP',
              'GaiaEDR3_RP', 'WISE_W1', 'WISE_W2', 'WISE_W3', 'WISE_W4')


def make_mist(ne=607, nm=140, nf=19, na=5, seed=0, frac_missing=0.02,
              frac_trunc=0.10, noise=1e-3):
    """Dense synthetic MIST grid.

    Axes follow SURVEY.md section 8d: EEP 202.. step 1, mass =
    unique(round(geomspace(0.25, 30, nm), 2)), [Fe/H] -4.0..0.5, [a/Fe]
    -0.2..0.6.  Returns a dict with

    ``eep, mass, feh, afe``  f64 axis vectors (strictly increasing)
    ``values``               f64 ``[na, nf, nm, ne, 9]`` in MIST_COLUMNS order
    ``valid``                bool ``[na, nf, nm, ne]``; False = row absent from
                             the track tables (whole track missing, or the
                             track stops at a lower EEP)
    """
    rng = np.random.default_rng(seed)
    eep = 202.0 + np.arange(ne, dtype=np.float64)
    mass = np.unique(np.round(np.geomspace(0.25, 30.0, nm), 2))
    nm = len(mass)
    feh = np.round(np.linspace(-4.0, 0.5, nf), 2)
    afe = np.round(np.linspace(-0.2, 0.6, na), 2)

    A = afe[:, None, None, None]
    F = feh[None, :, None, None]
    M = mass[None, None, :, None]
    s = ((eep - eep[0]) / max(eep[-1] - eep[0], 1.0))[None, None, None, :]
    lm = np.log10(M)

    shape = (na, nf, nm, ne)
    v = np.empty(shape + (9,), dtype=np.float64)
    # smooth, physically ranged functions of the four labels
    log_teff = 3.76 + 0.30 * np.tanh(1.6 * lm) - 0.012 * F + 0.004 * A \
        - 0.22 * s ** 3 + 0.03 * np.sin(6.0 * s) * (1.0 - s)
    log_teff = np.clip(log_teff, 3.5, 4.2)
    log_r = 0.85 * lm - 0.02 * F + 0.25 * s + 1.9 * s ** 6
    log_l = 2.0 * log_r + 4.0 * (log_teff - np.log10(5770.0))
    log_g = 4.438 + lm - 2.0 * log_r
    log_age = 10.0 - 2.6 * lm + 0.05 * F + 0.9 * np.log10(0.02 + s) - 0.9 * np.log10(1.02)
    star_mass = M * (1.0 - 0.05 * s ** 4) + 0.0 * F
    surf_feh = F - 0.06 * s * (1.0 - s) * 4.0 + 0.0 * M
    surf_afe = A + 0.01 * s + 0.0 * F * M
    agewgt = 0.02 + 0.98 * (0.5 + 0.5 * np.cos(5.0 * s + lm)) ** 2 + 0.0 * F * A
    for c, arr in enumerate((log_age, star_mass, log_r, log_l, log_teff,
                             surf_feh, surf_afe, log_g, agewgt)):
        v[..., c] = np.broadcast_to(arr, shape)
    v[..., :8] += noise * rng.standard_normal(shape + (8,))
    v[..., 8] = np.clip(v[..., 8] + noise * rng.standard_normal(shape), 1e-3, 1.0)

    valid = np.ones(shape, dtype=bool)
    ntrk = na * nf * nm
    trk = rng.random(ntrk).reshape(na, nf, nm)
    missing = trk < frac_missing
    trunc = (trk >= frac_missing) & (trk < frac_missing + frac_trunc)
    cut = rng.integers(max(ne // 4, 1), max(ne - 1, 2), size=(na, nf, nm))
    valid[missing] = False
    ee = np.arange(ne)[None, None, None, :]
    valid &= ~(trunc[..., None] & (ee >= cut[..., None]))
    return dict(eep=eep, mass=mass, feh=feh, afe=afe, values=v, valid=valid)


def mist_rows(mist):
    """Dense grid -> HDF5-like ``{'index': [...], key: structured rows}``.

    Keys are ``'{feh:4.2f}/{afe:4.2f}/0.00'`` (format from reference
    pullspectra.py:231); rows inside a key are ordered by mass then EEP.  This
    is the layout ``GenMIST.make_lib`` concatenates (reference
    fastMISTmod.py:76-98).
    """
    dt = np.dtype([(n, 'f8') for n in MIST_LABEL_FIELDS + MIST_COLUMNS])
    out = {'index': []}
    for j, f in enumerate(mist['feh']):
        for i, a in enumerate(mist['afe']):
            ok = mist['valid'][i, j]                     # [nm, ne]
            mi, ei = np.nonzero(ok)
            if len(mi) == 0:
                continue
            rows = np.empty(len(mi), dtype=dt)
            rows['EEP'] = mist['eep'][ei]
            rows['initial_mass'] = mist['mass'][mi]
            rows['initial_[Fe/H]'] = f
            rows['initial_[a/Fe]'] = a
            vals = mist['values'][i, j][mi, ei]          # [n, 9]
            for c, name in enumerate(MIST_COLUMNS):
                rows[name] = vals[:, c]
            key = '{0:4.2f}/{1:4.2f}/0.00'.format(f, a)
            out['index'].append(key)
            out[key] = rows
    return out


def make_phot_net(bands=DEMO_BANDS, h=128, seed=1, d_in=6):
    """Per-band MLP stack ``d_in -> h -> h -> 1`` with sigmoid activations.

    Shapes follow the recollected ``FastPayneSEDPredict`` stacking (SURVEY.md
    Appendix B): ``w1[nb,h,d_in]``, ``w2[nb,h,h]``, ``w3[nb,1,h]`` in f32;
    inputs are ``(Teff, logg, feh, afe, Av, Rv)`` scaled by
    ``(x - xmin)/(xmax - xmin) - 0.5`` (box from reference prior.py:117-129).
    """
    rng = np.random.default_rng(seed)
    nb = len(bands)
    f32 = np.float32
    # larger first-layer gain so the BC surface has structure over the box
    w1 = (rng.standard_normal((nb, h, d_in)) * (2.0 / np.sqrt(d_in))).astype(f32)
    b1 = (0.3 * rng.standard_normal((nb, h))).astype(f32)
    w2 = (rng.standard_normal((nb, h, h)) * (2.0 / np.sqrt(h))).astype(f32)
    b2 = (0.3 * rng.standard_normal((nb, h))).astype(f32)
    w3 = (rng.standard_normal((nb, 1, h)) * (1.0 / np.sqrt(h))).astype(f32)
    b3 = (0.5 * rng.standard_normal((nb, 1))).astype(f32)
    xmin = np.array([2500.0, -1.0, -4.0, -0.2, 0.0, 2.0])[:d_in]
    xmax = np.array([15000.0, 5.5, 0.5, 0.6, 5.0, 5.0])[:d_in]
    return dict(bands=np.array(bands), w1=w1, b1=b1, w2=w2, b2=b2, w3=w3, b3=b3,
                xmin=xmin, xmax=xmax, activation=np.array('sigmoid'))


def make_spec_net(npix=10240, h=300, d_in=4, seed=2, wave0=5140.0,
                  r_nn=100000.0, nlines=900, leak=0.01, logt_input=False):
    """Spectral MLP ``d_in -> h -> h -> npix`` (leaky-ReLU), Payne YST shape.

    The wavelength grid is log-uniform with three pixels per resolution
    element, ``dlnlam = 1/(3 r_nn)`` (reference pullspectra.py:322-330).
    ``resolution`` is the value the predictor hands to
    ``smoothspec(..., inres=resolution)``.  The output
    layer is built from absorption-line profiles so the predicted flux looks
    like a continuum-normalised spectrum in (0, 1.1].
    """
    rng = np.random.default_rng(seed)
    f32 = np.float32
    w0 = (rng.standard_normal((h, d_in)) * (1.5 / np.sqrt(d_in))).astype(f32)
    b0 = (0.2 * rng.standard_normal(h)).astype(f32)
    w1 = (rng.standard_normal((h, h)) * (1.2 / np.sqrt(h))).astype(f32)
    b1 = (0.2 * rng.standard_normal(h)).astype(f32)
    pix = np.arange(npix, dtype=np.float64)
    centers = rng.uniform(0, npix, nlines)
    depth = rng.uniform(0.02, 0.6, nlines) ** 1.5
    width = rng.uniform(1.3, 2.2, nlines)                  # sigma in pixels
    load = np.abs(rng.standard_normal((nlines, h))) / h      # hidden loadings
    w2 = np.zeros((npix, h), dtype=np.float64)
    b2 = np.ones(npix, dtype=np.float64)
    for c, d, s, u in zip(centers, depth, width, load):
        lo, hi = int(max(c - 6 * s, 0)), int(min(c + 6 * s + 1, npix))
        prof = d * np.exp(-0.5 * ((pix[lo:hi] - c) / s) ** 2)
        w2[lo:hi] -= 0.6 * prof[:, None] * u[None, :] * 4.0
        b2[lo:hi] -= 0.4 * prof
    w2 += 0.002 * rng.standard_normal((npix, h)) / np.sqrt(h)
    wavelength = wave0 * np.exp(pix / (3.0 * r_nn))
    xmin = np.array([2500.0, -1.0, -4.0, -0.2, 0.5])[:d_in]
    xmax = np.array([15000.0, 5.5, 0.5, 0.6, 3.0])[:d_in]
    if logt_input:
        xmin[0], xmax[0] = np.log10(xmin[0]), np.log10(xmax[0])
    return dict(w0=w0, b0=b0, w1=w1, b1=b1, w2=w2.astype(f32), b2=b2.astype(f32),
                xmin=xmin, xmax=xmax, wavelength=wavelength,
                resolution=np.float64(r_nn), leak=np.float64(leak),
                logt_input=np.bool_(logt_input))


def demo_obs_phot(bands=DEMO_BANDS):
    """The demo star's observed magnitudes with the demo's noise floors
    (reference demo/phot/18Sco_phot.dat, demo/runMS.py:68-85)."""
    raw = {'2MASS_J': (4.667, 0.260), '2MASS_H': (4.162, 0.178),
           '2MASS_Ks': (4.186, 0.292), 'GaiaEDR3_G': (5.327651, 0.002811),
           'GaiaEDR3_BP': (5.654931, 0.002940), 'GaiaEDR3_RP': (4.837854, 0.005691),
           'WISE_W1': (3.988, 0.234), 'WISE_W2': (3.642, 0.173),
           'WISE_W3': (3.994, 0.013), 'WISE_W4': (3.956, 0.024)}
    out = {}
    for b in bands:
        m, e = raw[b]
        floor = 0.01 if b.startswith('GaiaEDR3') else 0.05
        out[b] = [m, float(np.sqrt(e ** 2 + floor ** 2))]
    return out


def make_obs_spectrum(spec_net, n_obs=1536, snr=150.0, seed=4,
                      wave_lo=5146.0, wave_hi=5300.0, flux=None):
    """Observed wavelength grid (linear, like the demo HARPS spectrum: 1536 px
    over 5143.6-5301.2 A) with a placeholder flux
    of 1 +- noise; callers overwrite ``obs_flux`` with a model at a truth
    point when they need a realistic chi-square surface."""
    rng = np.random.default_rng(seed)
    w = np.linspace(wave_lo, wave_hi, n_obs)
    f = np.ones(n_obs) if flux is None else np.asarray(flux, dtype=np.float64)
    e = np.full(n_obs, 1.0 / snr)
    f = f + e * rng.standard_normal(n_obs)
    return dict(obs_wave=w, obs_flux=f, obs_eflux=e)


def draw_theta_phot(n, seed=3, frac_out=0.01, box=None):
    """Config-2 live points: columns ``(Dist, Av, EEP, initial_[Fe/H],
    initial_[a/Fe], initial_Mass)`` uniform in the demo prior box (reference
    demo/runMS.py:131-137), with ``frac_out`` of them pushed out of the grid."""
    rng = np.random.default_rng(seed)
    box = box or dict(Dist=(1.0, 100.0), Av=(0.0, 0.1), EEP=(202.0, 808.0),
                      feh=(-4.0, 0.0), afe=(-0.2, 0.6), mass=(0.5, 1.25))
    cols = [rng.uniform(*box[k], size=n) for k in ('Dist', 'Av', 'EEP', 'feh', 'afe', 'mass')]
    th = np.stack(cols, axis=1)
    nout = int(frac_out * n)
    if nout:
        idx = rng.choice(n, nout, replace=False)
        which = rng.integers(0, 4, nout)
        th[idx[which == 0], 2] = 808.0 + rng.uniform(0, 5, (which == 0).sum())   # EEP above
        th[idx[which == 1], 5] = 0.2                                              # mass below
        th[idx[which == 2], 3] = 0.5 + rng.uniform(0, 1, (which == 2).sum())      # feh at/above top
        th[idx[which == 3], 4] = -0.3                                             # afe below
    return th
