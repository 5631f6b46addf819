"""CPU oracle for the MINESweeper likelihood hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and there only as the checker (or as the CPU
arm being timed), never as a fallback for the CUDA path.

Contents
--------
``refshim``         imports the reference's *own* modules from /root/reference with
                    stubbed h5py / astropy / dynesty / Payne (this container only;
                    used to pin the ports below and to write tests/golden/*.npz)
``mist_port``       NumPy restatement of reference minesweeper/fastMISTmod.py
``smoothing_port``  NumPy restatement of the FFT branches of reference
                    minesweeper/smoothing.py used on the path ('R' and 'vsini')
``payne_port``      NumPy restatement of the three entry points of the external,
                    un-vendored ``Payne`` package -- PARITY UNPINNED (see header there)
``like_port``       NumPy restatement of reference minesweeper/likelihood.py +
                    genmod.py + fitutils.polycalc (per-point, f64)

Pinning status: ``mist_port``, ``smoothing_port`` and ``like_port`` are checked
against outputs of the reference itself (tests/golden/*.npz, written by
tests/golden/make_golden.py through ``refshim``).  ``payne_port`` has no
reference source to check against: PARITY UNPINNED for the ANN arithmetic.
"""
