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
