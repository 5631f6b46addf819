"""ORACLE (test infrastructure only) -- times the reference's CPU execution model
of the likelihood hot path on the host cores: the per-point Python loop
``lnlikefn(theta_i)`` (reference fitstar.py:715-727 -> likelihood.py:92-181), one
worker process per core (the reference's intended scale-out is one OS process
per star, fitstar.py:244-249), BLAS threads pinned to 1.

Used only by bench.py (``cpu_baseline`` leg and ``--impl reference``).  The
likelihood object is the reference's OWN ``minesweeper.likelihood.likelihood`` class
(``kind: "reference"``) imported through oracle/refshim.py from /root/reference or, on
the GPU box, from the verbatim copy ``oracle/_ref`` that oracle/make_ref.py writes
(un-vendored Payne / h5py stubbed as for the golden vectors).  Only when neither is
present does it fall back to oracle/like_port.py (``kind: "port"``).

Must be called BEFORE the parent process initialises CUDA (workers are forked).
"""
import multiprocessing as mp
import os
import time

import numpy as np

_PORT = None
_THETA = None


def _work(args):
    lo, hi, t_start = args
    L, th = _PORT, _THETA
    while time.time() < t_start:          # common start line
        pass
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(lo, hi):
        v = L.lnlikefn(list(th[i % len(th)]))
        if v == v and v != -np.inf:
            acc += v
    return hi - lo, time.perf_counter() - t0, acc


def make_likeobj(fitargs, fitpars, runbools, kind=None):
    """(likelihood object, kind).  kind None = the reference's own class when importable."""
    from . import refshim
    if kind is None:
        kind = 'reference' if refshim.available() else 'port'
    if kind == 'port':
        from .like_port import LikelihoodPort
        return LikelihoodPort(fitargs, fitpars, runbools), 'port'
    import contextlib
    import io
    from .like_port import _mist_rows
    refshim.install()
    from minesweeper.likelihood import likelihood as RefLikelihood
    fa = dict(fitargs)
    if fitpars[1].get('EEP', False):
        refshim.register_h5('cpuarm://mist', _mist_rows(fitargs['MISTpath']))
        fa['MISTpath'] = 'cpuarm://mist'
    with contextlib.redirect_stdout(io.StringIO()):
        obj = RefLikelihood(fa, fitpars, runbools, verbose=False)
    return obj, 'reference'


class CpuArm(object):
    def __init__(self, fitargs, fitpars, runbools, theta, cores=None, kind=None):
        global _PORT, _THETA
        for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
            os.environ.setdefault(k, '1')
        self.cores = cores or (len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity')
                               else os.cpu_count())
        _PORT, self.kind = make_likeobj(fitargs, fitpars, runbools, kind)
        _THETA = np.asarray(theta, dtype=np.float64)
        self.pool = mp.get_context('fork').Pool(self.cores)
        self.rate1 = None

    def calibrate(self, seconds=0.5):
        """single-core rate (evals/s) from a short probe in the parent."""
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            _PORT.lnlikefn(list(_THETA[n % len(_THETA)]))
            n += 1
        self.rate1 = n / (time.perf_counter() - t0)
        return self.rate1

    def eval_rows(self, rows):
        """lnL of theta[rows] evaluated in this process by the same likelihood object (parity reference)."""
        with np.errstate(all='ignore'):
            return np.array([_PORT.lnlikefn(list(_THETA[i])) for i in rows])

    def step(self, n_total):
        """n_total evaluations split over the workers; returns (evals, wall seconds)."""
        per = max(n_total // self.cores, 1)
        t_start = time.time() + 0.05
        jobs = [(w * per, (w + 1) * per, t_start) for w in range(self.cores)]
        t0 = time.perf_counter()
        res = self.pool.map(_work, jobs)
        wall = time.perf_counter() - t0 - 0.05
        return sum(r[0] for r in res), max(max(r[1] for r in res), wall)

    def close(self):
        self.pool.close()
        self.pool.join()
