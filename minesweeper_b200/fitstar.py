"""The caller side of the hot path (reference minesweeper/fitstar.py:715-727):
``lnprobfn`` with the reference's signature, plus the vectorised form a batched
sampler uses.  The prior object is the reference's own (scalar, CPU) ``prior``;
it consumes the per-row ``parsdict`` the likelihood produces.
"""
import numpy as np


def lnprobfn(pars, likeobj, priorobj):
    lnlike = likeobj.lnlikefn(pars)
    if lnlike == -np.inf:
        return -np.inf
    lnprior = priorobj.lnpriorfn(likeobj.parsdict)
    if lnprior == -np.inf:
        return -np.inf
    return lnprior + lnlike


def lnprobfn_batch(theta, likeobj, priorobj):
    """theta[B, ndim] -> lnprob[B]; one kernel sequence for the whole batch, then
    the (still scalar) prior per surviving row."""
    lnl, dicts = likeobj.lnlikefn_batch(theta)
    out = np.full(len(lnl), -np.inf)
    for i, (l, pd) in enumerate(zip(lnl, dicts)):
        if l == -np.inf:
            continue
        lp = priorobj.lnpriorfn(pd)
        if lp == -np.inf:
            continue
        out[i] = lp + l
    if len(dicts):
        likeobj.parsdict = dicts[-1]
    return out


class BatchPool(object):
    """``pool``-like object for dynesty (``pool=BatchPool(...), queue_size=N``):
    ``map(fn, iterable)`` evaluates all queued points in one batched call when
    ``fn`` is dynesty's wrapped log-likelihood over ``lnprobfn``; anything else
    is mapped serially."""

    def __init__(self, likeobj, priorobj, size=1024):
        self.likeobj, self.priorobj, self.size = likeobj, priorobj, size

    def map(self, fn, iterable):
        items = list(iterable)
        target = getattr(fn, 'func', None) or getattr(fn, 'loglikelihood', None) or fn
        if target is lnprobfn or getattr(target, '__name__', '') == 'lnprobfn':
            return list(lnprobfn_batch(np.asarray(items, dtype=np.float64), self.likeobj, self.priorobj))
        return [fn(x) for x in items]
