"""Property tests (hypothesis) of the interpolator semantics on the CPU oracle, and of the CUDA
brackets against np.digitize on the GPU."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import mist_port
from minesweeper_b200 import synth

MIST = synth.make_mist(ne=16, nm=9, nf=5, na=4, seed=77, frac_missing=0.1, frac_trunc=0.3)
AX = [MIST['eep'], MIST['mass'], MIST['feh'], MIST['afe']]
fl = lambda a: st.floats(min_value=float(a[0]) - 1.0, max_value=float(a[-1]) + 1.0, allow_nan=False)


@settings(max_examples=200, deadline=None)
@given(fl(AX[0]), fl(AX[1]), fl(AX[2]), fl(AX[3]))
def test_dense_oracle_properties(e, m, f, a):
    br, pred, ok = mist_port.interp_dense(MIST, [e], [m], [f], [a])
    want = [np.digitize([v], ax, right=False)[0] - 1 for v, ax in zip((e, m, f, a), AX)]
    assert list(br[0]) == want                                   # digitize semantics on every axis
    inside = all(0 <= b <= len(ax) - 2 for b, ax in zip(want, AX))
    if not inside:
        assert not ok[0] and np.isnan(pred[0]).all()             # out of grid (incl. the top value) -> None
    if ok[0]:
        # convex combination of the corner values: inside the corner range for every column
        i = [int(b) for b in br[0]]
        corners = MIST['values'][i[3]:i[3] + 2, i[2]:i[2] + 2, i[1]:i[1] + 2, i[0]:i[0] + 2].reshape(-1, 9)
        valid = MIST['valid'][i[3]:i[3] + 2, i[2]:i[2] + 2, i[1]:i[1] + 2, i[0]:i[0] + 2].reshape(-1)
        lo, hi = corners[valid].min(0), corners[valid].max(0)
        assert np.all(pred[0] >= lo - 1e-9) and np.all(pred[0] <= hi + 1e-9)


def test_exact_grid_hit_returns_vertex():
    idx = np.argwhere(MIST['valid'][:-1, :-1, :-1, :-1])
    for ia, jf, im, ie in idx[::17]:
        _, pred, ok = mist_port.interp_dense(MIST, [AX[0][ie]], [AX[1][im]], [AX[2][jf]], [AX[3][ia]])
        assert ok[0]
        np.testing.assert_allclose(pred[0], MIST['values'][ia, jf, im, ie], rtol=1e-14)


@pytest.mark.gpu
def test_cuda_brackets_equal_digitize_everywhere():
    from minesweeper_b200.engine import Engine
    rng = np.random.default_rng(5)
    n = 200000
    q = np.stack([rng.uniform(ax[0] - 1, ax[-1] + 1, n) for ax in AX], 1)
    # sprinkle exact grid values, their neighbours in floating point, infinities and NaNs
    for d, ax in enumerate(AX):
        k = rng.integers(0, n, 4000)
        v = rng.choice(ax, 4000)
        q[k, d] = np.where(rng.random(4000) < 0.5, v, np.nextafter(v, rng.choice([-np.inf, np.inf], 4000)))
        q[rng.integers(0, n, 50), d] = np.nan
        q[rng.integers(0, n, 50), d] = np.inf
        q[rng.integers(0, n, 50), d] = -np.inf
    eng = Engine(mist=MIST, precision='f32')
    br, der, ok = eng.mist_interp(q)
    want = np.stack([np.digitize(q[:, d], AX[d], right=False) - 1 for d in range(4)], 1)
    assert np.array_equal(br.cpu().numpy(), want)
    rbr, rpred, rok = mist_port.interp_dense(MIST, q[:, 0], q[:, 1], q[:, 2], q[:, 3])
    assert np.array_equal(ok.cpu().numpy().astype(bool), rok)
    np.testing.assert_allclose(der.cpu().numpy()[rok][:, 4:], rpred[rok], rtol=3e-6, atol=3e-6)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10 ** 6), st.floats(0.0, 0.4), st.floats(0.0, 0.5))
def test_ingest_roundtrip_with_missing_and_truncated_tracks(seed, frac_missing, frac_trunc):
    """ingest.mist_dense_from_tables inverts synth.mist_rows for any pattern of missing / truncated tracks:
    same axes (as long as every label value survives somewhere), same validity mask, same values; table row
    order does not matter."""
    from minesweeper_b200.ingest import mist_dense_from_tables
    mist = synth.make_mist(ne=12, nm=6, nf=4, na=3, seed=seed, frac_missing=frac_missing, frac_trunc=frac_trunc)
    tables = synth.mist_rows(mist)
    if not tables['index']:
        return
    rng = np.random.default_rng(seed)
    for k in tables['index']:
        tables[k] = tables[k][rng.permutation(len(tables[k]))]
    rng.shuffle(tables['index'])
    dense = mist_dense_from_tables(tables)
    # axes present in the tables are a subset of the generating axes; map dense back onto the full grid
    sel = [np.searchsorted(mist[k], dense[k]) for k in ('eep', 'mass', 'feh', 'afe')]
    for k, s_ in zip(('eep', 'mass', 'feh', 'afe'), sel):
        np.testing.assert_array_equal(mist[k][s_], dense[k])
    sub_valid = mist['valid'][np.ix_(sel[3], sel[2], sel[1], sel[0])]
    sub_vals = mist['values'][np.ix_(sel[3], sel[2], sel[1], sel[0])]
    np.testing.assert_array_equal(dense['valid'], sub_valid)
    np.testing.assert_array_equal(dense['values'][dense['valid']], sub_vals[sub_valid])
    assert dense['valid'].sum() == mist['valid'].sum()
