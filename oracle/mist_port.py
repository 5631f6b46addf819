"""ORACLE (test infrastructure only) -- NumPy f64 restatement of the MIST
mass-track interpolator, reference minesweeper/fastMISTmod.py.

Two forms:

``GenMISTPort``       per-point, same execution model as the reference: row
                      table + integer pixel coordinates + cKDTree ball query +
                      N-linear weights + renormalisation
                      (fastMISTmod.py:28-74 ctor, :76-98 make_lib, :110-125
                      lib_as_grid, :127-150 params_to_grid, :152-170 weights,
                      :172-192 knearest_inds, :194-213 linear_weights,
                      :100-108 getMIST).  This is what the CPU baseline times.
``interp_dense``      vectorised over a batch on the dense ``[na,nf,nm,ne,9]``
                      layout; visits exactly the <= 16 corners
                      ``{i_d, i_d+1}`` that can carry non-zero weight
                      (SURVEY.md Appendix A.4).  Used for parity checks at
                      sizes where the per-point form is too slow.

Both are pinned against the reference module itself by tests/golden/mist.npz.
"""
import numpy as np
from scipy.spatial import cKDTree

LABELS = ('EEP', 'initial_Mass', 'initial_[Fe/H]', 'initial_[a/Fe]')
_ROW_FIELDS = ('EEP', 'initial_mass', 'initial_[Fe/H]', 'initial_[a/Fe]')
_OUT_FIELDS = ('log_age', 'star_mass', 'log_R', 'log_L', 'log_Teff',
               '[Fe/H]', '[a/Fe]', 'log_g')
PREDICTIONS = ('log(Age)', 'Mass', 'log(R)', 'log(L)', 'log(Teff)', '[Fe/H]',
               '[a/Fe]', 'log(g)')


def dense_to_rows(mist):
    """Dense container ``{eep, mass, feh, afe, values[na,nf,nm,ne,9], valid}`` -> the HDF5-like
    ``{'index': [...], key: structured rows}`` that ``GenMIST.make_lib`` concatenates
    (fastMISTmod.py:76-98): one table per ([Fe/H], [a/Fe]) pair keyed
    ``'{feh:4.2f}/{afe:4.2f}/0.00'`` (key format of reference pullspectra.py:231), rows ordered by
    mass then EEP.  The oracle's own converter, so that the checker does not depend on product code."""
    dt = np.dtype([(n, 'f8') for n in _ROW_FIELDS + _OUT_FIELDS + ('Agewgt',)])
    out = {'index': []}
    for j, f in enumerate(mist['feh']):
        for i, a in enumerate(mist['afe']):
            mi, ei = np.nonzero(mist['valid'][i, j])
            if len(mi) == 0:
                continue
            rows = np.empty(len(mi), dtype=dt)
            rows['EEP'] = mist['eep'][ei]
            rows['initial_mass'] = mist['mass'][mi]
            rows['initial_[Fe/H]'] = f
            rows['initial_[a/Fe]'] = a
            vals = mist['values'][i, j][mi, ei]
            for c, name in enumerate(_OUT_FIELDS + ('Agewgt',)):
                rows[name] = vals[:, c]
            key = '{0:4.2f}/{1:4.2f}/0.00'.format(f, a)
            out['index'].append(key)
            out[key] = rows
    return out


def bracket(axis, x):
    """``digitize([x], axis, right=False) - 1``: the i with
    ``axis[i] <= x < axis[i+1]``; -1 below the axis, ``len(axis)-1`` at or
    above its top value and for NaN (fastMISTmod.py:138)."""
    return np.searchsorted(axis, x, side='right') - 1


class GenMISTPort(object):
    def __init__(self, rows, ageweight=True):
        keys = list(rows['index'])
        lab = np.stack([np.concatenate([np.asarray(rows[k][f], dtype=np.float64) for k in keys])
                        for f in _ROW_FIELDS], axis=1)
        # the three label columns that are rounded to two decimals (:89-91)
        lab[:, 1:] = np.around(lab[:, 1:], decimals=2)
        fields = _OUT_FIELDS + (('Agewgt',) if ageweight else ())
        self.output = np.stack([np.concatenate([np.asarray(rows[k][f], dtype=np.float64) for k in keys])
                                for f in fields], axis=1)
        self.modpararr = list(LABELS) + list(PREDICTIONS) + (['Agewgt'] if ageweight else [])
        self.axes = [np.unique(lab[:, d]) for d in range(4)]
        self.widths = [np.diff(a) for a in self.axes]
        # integer pixel coordinate of every row (:121-123)
        self.X = np.stack([np.searchsorted(self.axes[d], lab[:, d], side='left')
                           for d in range(4)], axis=1)
        self.tree = cKDTree(self.X)
        self.radius = np.sqrt(4 * 2.0 ** 2)            # :69-74

    def pixel_coords(self, eep, mass, feh, afe):
        """(brackets[4], xtarg[4]); raises ValueError out of grid (:127-150)."""
        q = (eep, mass, feh, afe)
        idx = np.array([bracket(self.axes[d], q[d]) for d in range(4)])
        frac = np.empty(4)
        for d in range(4):
            i = idx[d]
            if i >= len(self.widths[d]):               # IndexError in the reference
                raise ValueError('outside grid')
            # i == -1 wraps to the last cell exactly like NumPy negative indexing
            frac[d] = (q[d] - self.axes[d][i]) / self.widths[d][i]
        return idx, idx + frac

    def corner_weights(self, eep, mass, feh, afe):
        _, xt = self.pixel_coords(eep, mass, feh, afe)
        cand = np.sort(self.tree.query_ball_point(xt.reshape(1, -1), self.radius)[0]).astype(np.int64)
        if len(cand) == 0:
            raise ValueError('no neighbours')
        dx = xt - self.X[cand]
        w = np.where(dx >= 0, 1.0 - dx, 1.0 + dx)
        w = w * ((dx > -1) & (dx < 1))
        w = w.prod(axis=-1)
        if w.sum() <= 0.0:
            raise ValueError('zero weight')
        keep = w > 0
        cand, w = cand[keep], w[keep]
        return cand, w / w.sum()

    def getMIST(self, mass=1.0, eep=300, feh=0.0, afe=0.0, **kw):
        try:
            rows, w = self.corner_weights(eep, mass, feh, afe)
        except ValueError:
            return None
        return [eep, mass, feh, afe] + list(np.dot(w, self.output[rows]))


def interp_dense(mist, eep, mass, feh, afe):
    """Batched restatement on the dense layout.

    Returns ``(brackets[B,4] int32, pred[B,9] f64, ok[B] bool)``.  ``brackets``
    equals ``digitize - 1`` per axis for every row, in or out of grid.  Rows
    that the reference answers with ``None`` have ``ok=False`` and NaN
    predictions.
    """
    axes = (mist['eep'], mist['mass'], mist['feh'], mist['afe'])
    q = [np.asarray(v, dtype=np.float64) for v in (eep, mass, feh, afe)]
    B = len(q[0])
    vals, valid = mist['values'], mist['valid']
    idx = np.stack([bracket(axes[d], q[d]) for d in range(4)], axis=1).astype(np.int32)
    inside = np.ones(B, dtype=bool)
    for d in range(4):
        inside &= (idx[:, d] >= 0) & (idx[:, d] <= len(axes[d]) - 2)
    lo = np.empty((B, 4))
    hi = np.empty((B, 4))
    ic = np.zeros((B, 4), dtype=np.int64)
    for d in range(4):
        i = np.where(inside, idx[:, d], 0)
        ic[:, d] = i
        frac = (q[d] - axes[d][i]) / (axes[d][i + 1] - axes[d][i])
        xt = i + frac
        dl = xt - i
        du = xt - (i + 1)
        lo[:, d] = 1.0 - dl
        hi[:, d] = np.where(du > -1, 1.0 + du, 0.0)
    pred = np.zeros((B, vals.shape[-1]))
    wsum = np.zeros(B)
    # ascending row id in the reference = (feh, afe, mass, eep) nesting of the
    # track tables; visit corners in that order
    for cf in (0, 1):
        for ca in (0, 1):
            for cm in (0, 1):
                for ce in (0, 1):
                    w = (hi[:, 0] if ce else lo[:, 0])
                    w = w * (hi[:, 1] if cm else lo[:, 1])
                    w = w * (hi[:, 2] if cf else lo[:, 2])
                    w = w * (hi[:, 3] if ca else lo[:, 3])
                    ie, im = ic[:, 0] + ce, ic[:, 1] + cm
                    jf, ja = ic[:, 2] + cf, ic[:, 3] + ca
                    here = valid[ja, jf, im, ie] & (w > 0) & inside
                    w = np.where(here, w, 0.0)
                    wsum += w
                    pred += w[:, None] * np.where(here[:, None], vals[ja, jf, im, ie], 0.0)
    ok = inside & (wsum > 0)
    with np.errstate(invalid='ignore', divide='ignore'):
        pred = np.where(ok[:, None], pred / wsum[:, None], np.nan)
    return idx, pred, ok
