"""CPU: the oracle ports against golden vectors written by the reference itself
(tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import fit_setup, sub
from oracle import like_port, mist_port, smoothing_port
from minesweeper_b200 import synth


def test_mist_port_matches_reference(g_mist):
    mist = sub(g_mist, 'mist_')
    G = mist_port.GenMISTPort(synth.mist_rows(mist), ageweight=True)
    q = g_mist['theta4']
    step = 7                                   # per-point port is slow; the dense form below covers all rows
    for i in range(0, len(q), step):
        out = G.getMIST(eep=q[i, 0], mass=q[i, 1], feh=q[i, 2], afe=q[i, 3])
        if g_mist['ok'][i]:
            np.testing.assert_allclose(out, g_mist['pred'][i], rtol=1e-14, atol=0)
        else:
            assert out is None


def test_mist_dense_matches_reference(g_mist):
    mist = sub(g_mist, 'mist_')
    q = g_mist['theta4']
    br, pred, ok = mist_port.interp_dense(mist, q[:, 0], q[:, 1], q[:, 2], q[:, 3])
    assert np.array_equal(br, g_mist['brackets'])              # bit-exact brackets, in and out of grid
    assert np.array_equal(ok, g_mist['ok'])
    np.testing.assert_allclose(pred[ok], g_mist['pred'][ok][:, 4:], rtol=1e-13, atol=1e-15)
    assert g_mist['ok'].sum() > 2000 and (~g_mist['ok']).sum() > 100


def test_smoothing_port_matches_reference(g_smooth):
    w, f, ow = g_smooth['wave'], g_smooth['flux'], g_smooth['outwave']
    for v, ref in zip(g_smooth['vsini'], g_smooth['vsini_out']):
        out = smoothing_port.smooth_vsini(w, f, v, outwave=None, inres=0.0)
        np.testing.assert_allclose(out, ref, rtol=0, atol=1e-14)
    ws = w * (1 + float(g_smooth['r_vshift']) / 299792.458)
    for r, ref in zip(g_smooth['rsigma'], g_smooth['r_out']):
        out = smoothing_port.smooth_R(ws, f, r, ow, inres=float(g_smooth['r_inres']))
        np.testing.assert_allclose(out, ref, rtol=0, atol=1e-14)


def test_lsf_smoothing_port_matches_reference(g_smooth):
    """Row a14 (array Inst_R / wavelength-dependent LSF, smoothing.py:412-514): oracle restatement against the
    reference's own output (the CUDA form is checked against the same vectors in test_gpu_parity.py)."""
    w, f, ow = g_smooth['wave'], g_smooth['flux'], g_smooth['outwave']
    for (a, b), ref in zip(g_smooth['lsf_ab'], g_smooth['lsf_out']):
        out = smoothing_port.smooth_lsf(w, f, ow, a + b * (w - w[0]))
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('which', ['phot', 'spec'])
def test_like_port_matches_reference(which, g_like_phot, g_like_spec):
    g = g_like_phot if which == 'phot' else g_like_spec
    fitargs, fitpars, runbools = fit_setup(g)
    L = like_port.LikelihoodPort(fitargs, fitpars, runbools)
    assert L.fitpars_i == [str(s) for s in g['fitpars_i']]
    keys = [str(k) for k in g['parsdict_keys']]
    n = len(g['theta']) if which == 'phot' else 40
    for i in range(n):
        lnl = L.lnlikefn(list(g['theta'][i]))
        if np.isfinite(g['lnl'][i]):
            assert abs(lnl - g['lnl'][i]) <= 1e-9 * max(1.0, abs(g['lnl'][i])), i
            got = np.array([L.parsdict.get(k, np.nan) for k in keys])
            np.testing.assert_allclose(got, g['parsdict'][i], rtol=1e-13, atol=0)
        else:
            assert lnl == g['lnl'][i]
            assert all(L.parsdict[k] == np.inf for k in like_port.FAIL_KEYS)
    assert (~np.isfinite(g['lnl'])).sum() >= 2


def test_mist_ingest_roundtrip_and_reference_lib(tmp_path):
    """minesweeper_b200.ingest: HDF5-like tables -> dense grid reproduces the grid the tables came
    from, and agrees row by row with what the reference's make_lib / lib_as_grid build from the
    same tables (fastMISTmod.py:76-125)."""
    from minesweeper_b200 import synth
    from minesweeper_b200.ingest import mist_dense_from_tables
    mist = synth.make_mist(ne=30, nm=9, nf=4, na=3, seed=5)
    tables = synth.mist_rows(mist)
    dense = mist_dense_from_tables(tables)
    for k in ('eep', 'mass', 'feh', 'afe'):
        np.testing.assert_array_equal(dense[k], mist[k])
    np.testing.assert_array_equal(dense['valid'], mist['valid'])
    np.testing.assert_array_equal(dense['values'][dense['valid']], mist['values'][mist['valid']])
    from oracle import refshim
    if not refshim.available():
        return
    refshim.install()
    refshim.register_h5('ingest://mist', tables)
    from minesweeper.fastMISTmod import GenMIST
    G = GenMIST(MISTpath='ingest://mist', ageweight=True, verbose=False)
    ie, im, jf, ja = G.X.T
    np.testing.assert_array_equal(dense['values'][ja, jf, im, ie], G.output)
    assert dense['valid'].sum() == len(G.output)


def test_postproc_weighted_quantiles(tmp_path):
    """minesweeper_b200.postproc against the oracle's restatement of Payne's weighted quantile and against an
    analytic case (equal weights on a fine grid)."""
    from minesweeper_b200 import postproc
    from oracle.payne_port import quantile
    rng = np.random.default_rng(11)
    x = rng.normal(2.0, 0.5, 4000)
    w = rng.uniform(0.0, 1.0, 4000)
    q = [0.001, 0.16, 0.5, 0.84, 0.999]
    np.testing.assert_allclose(postproc.weighted_quantile(x, q, w), quantile(x, q, weights=w), rtol=0, atol=1e-14)
    g = np.linspace(0.0, 1.0, 100001)
    np.testing.assert_allclose(postproc.weighted_quantile(g, [0.16, 0.5, 0.84], np.ones_like(g)), [0.16, 0.5, 0.84],
                               atol=2e-5)
    # a samples file in the reference layout: quantiles of a weighted Gaussian posterior
    n = 3000
    logwt = -0.5 * ((x[:n] - 2.0) / 0.5) ** 2 * 0.0 + np.log(w[:n] + 1e-12)
    path = tmp_path / 'star_samp.dat'
    with open(path, 'w') as f:
        f.write('Iter Dist log(lk) log(vol) log(wt) h nc log(z) delta(log(z))\n')
        for i in range(n):
            f.write('%d %.12g 0 0 %.12g nan nan nan nan\n' % (i, x[i], logwt[i]))
    st = postproc.summarize_samples_file(str(path))['Dist']
    ref = quantile(x[:n], [0.16, 0.5, 0.84], weights=w[:n] + 1e-12)
    np.testing.assert_allclose(st[:3], ref, rtol=1e-9)
    assert abs(st[3] - 0.5) < 0.05 and st[4][0] < st[0] < st[2] < st[4][1]


def test_samples_file_roundtrip_both_record_layouts(tmp_path):
    """sampler.write_samples_file -> postproc.read_samples_file for the per-iteration tuples of the torch sampler
    and for the per-star arrays of the kernel-bookkeeping sampler: same file either way."""
    from minesweeper_b200 import postproc
    from minesweeper_b200.sampler import write_samples_file
    rng = np.random.default_rng(3)
    names = ['Dist', 'Av', 'EEP']
    n = 50
    rows = rng.normal(size=(n, 3))
    lnl, logvol, logwt = np.sort(rng.normal(size=n)), -np.arange(1, n + 1) / 150.0, rng.normal(size=n)
    per_star = [np.zeros((0, 6)), np.column_stack([rows, lnl, logvol, logwt])]
    tuples = [(np.array([0, 1]), np.stack([rng.normal(size=3), rows[i]]), np.array([0.0, lnl[i]]),
               np.array([0.0, logvol[i]]), np.array([0.0, logwt[i]])) for i in range(n)]
    pa, pb = str(tmp_path / 'a.dat'), str(tmp_path / 'b.dat')
    write_samples_file(pa, dict(names=names, dead=per_star), star=1)
    write_samples_file(pb, dict(names=names, dead=tuples), star=1)
    assert open(pa).read() == open(pb).read()
    tab = postproc.read_samples_file(pa)
    np.testing.assert_allclose(tab['EEP'], rows[:, 2], rtol=1e-9)
    np.testing.assert_allclose(tab['log(wt)'], logwt, rtol=1e-9)
    st = postproc.summarize_samples_file(pa)
    assert set(st) == set(names) and st['Dist'][0] <= st['Dist'][1] <= st['Dist'][2]


def test_dense_batched_lnlike_matches_reference_golden(g_like_phot):
    """oracle.like_port.lnlike_phot_dense (the checker bench.py and the full-grid GPU test use at sizes
    where the per-point form is too slow) against lnL / mags written by the reference's own likelihood
    object (tests/golden/like_phot.npz)."""
    from oracle.like_port import lnlike_phot_dense
    g = g_like_phot
    fitargs, fitpars, runbools = fit_setup(g)
    assert [str(s) for s in g['fitpars_i']] == ['Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass']
    r = lnlike_phot_dense(fitargs['MISTpath'], fitargs['photANNpath'], fitargs['obs_phot'], g['theta'])
    fin = np.isfinite(g['lnl'])
    assert np.array_equal(np.isneginf(r['lnl']), np.isneginf(g['lnl']))
    np.testing.assert_allclose(r['lnl'][fin], g['lnl'][fin], rtol=1e-11, atol=1e-8)
    assert [str(b) for b in g['bands']] == r['bands']
    np.testing.assert_allclose(r['mags'][fin], g['mags'][fin], rtol=1e-12, atol=1e-12)


def test_reference_copy_and_cpu_arm_object():
    """oracle/make_ref.py copies the reference modules byte for byte, and the CPU arm's likelihood object is the
    reference's own class (kind 'reference') agreeing with the port."""
    import hashlib
    from oracle import make_ref, refshim
    from oracle.cpu_baseline import make_likeobj
    from minesweeper_b200 import synth
    if not refshim.available():
        pytest.skip('no reference checkout and no oracle/_ref copy')
    dst = make_ref.make(verbose=False)
    if os.path.isdir('/root/reference/minesweeper'):
        for f in make_ref.FILES:
            a = hashlib.sha256(open(os.path.join('/root/reference/minesweeper', f), 'rb').read()).hexdigest()
            b = hashlib.sha256(open(os.path.join(dst, 'minesweeper', f), 'rb').read()).hexdigest()
            assert a == b, f
    mist = synth.make_mist(ne=30, nm=9, nf=4, na=3, seed=5)
    phot = synth.make_phot_net(h=16, seed=1)
    names = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'Av',
             'log(A)', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass']
    on = {'Dist', 'Av', 'EEP', 'initial_[Fe/H]', 'initial_[a/Fe]', 'initial_Mass'}
    fitpars = [names, {n: (n in on) for n in names}]
    fitargs = dict(ageweight=True, mistdif=True, fixedpars={}, MISTpath=mist, photANNpath=phot,
                   obs_phot=synth.demo_obs_phot())
    rb = [False, True, False, False, False]
    R, kind = make_likeobj(fitargs, fitpars, rb)
    P, kp = make_likeobj(fitargs, fitpars, rb, 'port')
    assert kind == 'reference' and kp == 'port' and type(R).__module__ == 'minesweeper.likelihood'
    th = synth.draw_theta_phot(60, seed=3, frac_out=0.1,
                               box=dict(Dist=(1.0, 100.0), Av=(0.0, 0.1), EEP=(202.0, 231.0), feh=(-4.0, 0.5),
                                        afe=(-0.2, 0.6), mass=(0.25, 30.0)))
    a = np.array([R.lnlikefn(list(t)) for t in th])
    b = np.array([P.lnlikefn(list(t)) for t in th])
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize('which', ['phot', 'spec'])
def test_samples_file_header_is_the_references_parsdict_order(which, g_like_phot, g_like_spec):
    """The samples-file columns (sampler.samples_table with a likelihood object) follow the key order of the
    reference's own ``likeobj.parsdict`` after a successful call -- what ``fitstar._initoutput`` writes
    (fitstar.py:235-242, 392-402) and demo/postproc.py:142-160 reads."""
    from oracle import refshim
    from oracle.cpu_baseline import make_likeobj
    from minesweeper_b200.likelihood import parsdict_keys
    if not refshim.available():
        pytest.skip('no reference checkout and no oracle/_ref copy')
    g = g_like_phot if which == 'phot' else g_like_spec
    fitargs, fitpars, runbools = fit_setup(g)
    R, kind = make_likeobj(fitargs, fitpars, runbools)
    assert kind == 'reference'
    i = int(np.flatnonzero(np.isfinite(g['lnl']))[0])
    R.lnlikefn(list(g['theta'][i]))
    want = list(R.parsdict.keys())
    fitpars_i = [p for p in fitpars[0] if fitpars[1][p]]
    assert parsdict_keys(fitpars_i, fitargs['fixedpars'], iso=True, ageweight=True) == want


def test_payne_weight_file_ingestion_roundtrip():
    """minesweeper_b200.ingest: HDF5-like network files (layout of SURVEY.md Appendix B) -> containers that evaluate
    to the same numbers as the nets they were written from (through the oracle's CPU restatement)."""
    from minesweeper_b200 import ingest
    from oracle.payne_port import FastPayneSEDPredict, PayneSpecPredict
    phot = synth.make_phot_net(h=24, seed=5)
    files = {}
    for i, b in enumerate(phot['bands']):
        files[str(b)] = {'model': {'lin1.weight': phot['w1'][i], 'lin1.bias': phot['b1'][i], 'lin2.weight': phot['w2'][i],
                                   'lin2.bias': phot['b2'][i], 'lin3.weight': phot['w3'][i], 'lin3.bias': phot['b3'][i]},
                         'xmin': phot['xmin'], 'xmax': phot['xmax']}
    pick = ['WISE_W1', '2MASS_J', 'GaiaEDR3_G']
    c = ingest.payne_phot_net(files, bands=pick)
    assert [str(b) for b in c['bands']] == pick and c['w2'].shape == (3, 24, 24)
    x = [5600.0, 4.4, -0.3, 0.1, 0.05, 3.1]
    np.testing.assert_array_equal(FastPayneSEDPredict(nnpath=c).bolcorr(x),
                                  FastPayneSEDPredict(usebands=pick, nnpath=phot).bolcorr(x))
    spec = synth.make_spec_net(npix=512, h=16, seed=6, nlines=40)
    f = {'w_array_0': spec['w0'], 'w_array_1': spec['w1'], 'w_array_2': spec['w2'], 'b_array_0': spec['b0'],
         'b_array_1': spec['b1'], 'b_array_2': spec['b2'], 'x_min': spec['xmin'], 'x_max': spec['xmax'],
         'wavelength': spec['wavelength'], 'resolution': np.array([spec['resolution']])}
    cs = ingest.payne_spec_net(f)
    np.testing.assert_array_equal(PayneSpecPredict(nnpath=cs).predictspec(5600.0, 4.4, -0.3, 0.1),
                                  PayneSpecPredict(nnpath=spec).predictspec(5600.0, 4.4, -0.3, 0.1))
    # h5py is absent here (ImportError with the conversion recipe); when an earlier test installed the oracle's h5py
    # stub the lookup of the unknown file fails instead -- loudly either way
    with pytest.raises((ImportError, KeyError, OSError)):
        ingest.load_model_file('/nonexistent/MIST.h5', 'mist')


@pytest.mark.parametrize('which', ['phot', 'spec'])
def test_golden_lnl_is_what_the_reference_class_returns(which, g_like_phot, g_like_spec):
    """The committed vectors are outputs of the reference: its own ``likelihood.likelihood`` class (from the checkout
    or the verbatim copy in oracle/_ref), built from the model arrays stored in the golden file, returns the stored
    lnL and parsdict values again -- so a stale or hand-edited fixture cannot pass unnoticed."""
    from oracle import refshim
    from oracle.cpu_baseline import make_likeobj
    if not refshim.available():
        pytest.skip('no reference checkout and no oracle/_ref copy')
    g = g_like_phot if which == 'phot' else g_like_spec
    fitargs, fitpars, runbools = fit_setup(g)
    R, kind = make_likeobj(fitargs, fitpars, runbools)
    assert kind == 'reference' and type(R).__module__ == 'minesweeper.likelihood'
    keys = [str(k) for k in g['parsdict_keys']]
    n = 60 if which == 'phot' else 12
    import contextlib
    import io
    for i in range(n):
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
            lnl = R.lnlikefn(list(g['theta'][i]))
        if np.isfinite(g['lnl'][i]):
            assert abs(lnl - g['lnl'][i]) <= 1e-12 * max(1.0, abs(g['lnl'][i])), i
            got = np.array([R.parsdict.get(k, np.nan) for k in keys], dtype=np.float64)
            np.testing.assert_allclose(got, g['parsdict'][i], rtol=1e-14, atol=0)
        else:
            assert lnl == g['lnl'][i] or (np.isnan(lnl) and np.isnan(g['lnl'][i]))


def test_golden_smoothing_vectors_are_what_the_reference_module_returns(g_smooth):
    """tests/golden/smoothing.npz against the reference's own ``smoothing.smoothspec`` (checkout or oracle/_ref) on the
    stored spectrum: rotational, resolution (with Doppler shift and input resolution) and LSF-array modes."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip('no reference checkout and no oracle/_ref copy')
    refshim.install()
    from minesweeper import smoothing as S
    g = g_smooth
    w, f, ow = g['wave'], g['flux'], g['outwave']
    for v, ref in zip(g['vsini'], g['vsini_out']):
        out = S.smoothspec(w, f, v, outwave=None, smoothtype='vsini', fftsmooth=True, inres=0.0)
        np.testing.assert_array_equal(out, ref)
    ws = w * (1 + float(g['r_vshift']) / 299792.458)
    for r, ref in zip(g['rsigma'], g['r_out']):
        out = S.smoothspec(ws, f, r, outwave=ow, smoothtype='R', fftsmooth=True, inres=float(g['r_inres']))
        np.testing.assert_array_equal(out, ref)
    for (a, b), ref in zip(g['lsf_ab'], g['lsf_out']):
        out = S.smoothspec(w, f, a + b * (w - w[0]), outwave=ow, smoothtype='lsf', fftsmooth=True)
        np.testing.assert_array_equal(out, ref)
