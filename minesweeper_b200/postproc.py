"""Posterior summaries of a samples file: weighted 16/50/84 % quantiles and weighted standard deviation.

What ``demo/postproc.py`` derives from the dynesty output (reference demo/postproc.py:261-274, ``genstat``):
per parameter ``[q16, q50, q84, weighted std, plot range]`` with the weights ``exp(log(wt) - log(z)_final)``.
The weighted quantile is the one of ``Payne.utils.quantiles.quantile`` (external, un-vendored; restated in
``oracle/payne_port.py`` and therefore unpinned): sort, cumulative weights normalised to 1 with a leading 0,
linear interpolation.  Host-side NumPy: this is post-processing of one star's samples, not the hot path.
"""
import numpy as np


def weighted_quantile(x, q, weights=None):
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    q = np.atleast_1d(np.asarray(q, dtype=np.float64))
    if weights is None:
        return np.percentile(x, list(100.0 * q)).tolist()
    w = np.atleast_1d(np.asarray(weights, dtype=np.float64))
    order = np.argsort(x)
    cdf = np.cumsum(w[order])[:-1]
    cdf = np.append(0.0, cdf / cdf[-1])
    return np.interp(q, cdf, x[order]).tolist()


def read_samples_file(path):
    """Samples file (header fitstar.py:235-242) -> dict of columns."""
    with open(path) as f:
        names = f.readline().split()
        rows = np.array([[float(v) for v in line.split()] for line in f if line.strip()], dtype=np.float64)
    return {n: rows[:, i] for i, n in enumerate(names)}


def sample_weights(tab):
    """``exp(log(wt) - log(z)_final)``; when the file carries no running ln Z (our writer leaves it NaN) the final
    evidence is the log-sum of the ln weights."""
    logwt = tab['log(wt)']
    logz = tab.get('log(z)')
    if logz is not None and np.isfinite(logz[-1]):
        lz = logz[-1]
    else:
        m = np.max(logwt)
        lz = m + np.log(np.sum(np.exp(logwt - m)))
    return np.exp(logwt - lz)


def genstat(tab, pararr, wgt=None):
    """Per parameter: ``[q16, q50, q84, weighted std, [lo, hi]]`` as demo/postproc.py:261-274 builds it
    (including its asymmetric padding of the plot range: the upper pad uses the already padded lower edge)."""
    out = {}
    for kk in pararr:
        x = np.asarray(tab[kk], dtype=np.float64)
        st = weighted_quantile(x, [0.16, 0.5, 0.84], weights=wgt)
        avg = np.average(x, weights=wgt)
        st.append(float(np.sqrt(np.average((x - avg) ** 2.0, weights=wgt))))
        rng = weighted_quantile(x, [0.001, 0.999], weights=wgt)
        rng[0] = rng[0] - 0.15 * (rng[1] - rng[0])
        rng[1] = rng[1] + 0.15 * (rng[1] - rng[0])
        st.append(rng)
        out[kk] = st
    return out


def summarize_samples_file(path, pars=None):
    tab = read_samples_file(path)
    skip = {'Iter', 'log(lk)', 'log(vol)', 'log(wt)', 'h', 'nc', 'log(z)', 'delta(log(z))'}
    pars = [p for p in tab if p not in skip] if pars is None else list(pars)
    return genstat(tab, pars, wgt=sample_weights(tab))
