"""ORACLE (test infrastructure only) -- materialise ``oracle/_ref/``: the UNMODIFIED
reference modules of the likelihood path, taken from where they lie under
/root/reference, so that the CPU arm of bench.py (``--impl reference`` and the
``cpu_baseline`` leg) can time the reference's own code on the GPU box, where
/root/reference does not exist.

    python oracle/make_ref.py          (also run by __graft_entry__.build())

The reference is pure Python: "building" it is a verbatim file copy
plus the generated ``minesweeper/__init__.py`` that the reference's setup.py:31-43 would
write.  ``oracle/_ref/`` is git-ignored (no reference source enters the history) but not
gpurun-ignored, so it travels with a snapshot like the built ``.so``.  The un-vendored
``Payne`` package and ``h5py`` / ``astropy`` / ``dynesty`` are stubbed at import time by
oracle/refshim.py, exactly as for the golden vectors.

Files (SURVEY.md section 8a + the two callers): fastMISTmod, genmod, likelihood,
smoothing, fitutils, prior, advancedpriors (imported by prior), fitstar.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('MS_REFERENCE_ROOT', '/root/reference')
REF_DST = os.path.join(HERE, '_ref')
FILES = ('fastMISTmod.py', 'genmod.py', 'likelihood.py', 'smoothing.py', 'fitutils.py', 'prior.py',
         'advancedpriors.py', 'fitstar.py')


def make(verbose=True):
    """Copy the files; returns the destination, or None when the reference checkout is absent
    (GPU box: the prebuilt ``oracle/_ref`` of the snapshot is used as is)."""
    src = os.path.join(REF_SRC, 'minesweeper')
    if not os.path.isdir(src) or os.path.realpath(REF_SRC) == os.path.realpath(REF_DST):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, 'minesweeper')) else None
    dst = os.path.join(REF_DST, 'minesweeper')
    os.makedirs(dst, exist_ok=True)
    manifest = []
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        with open(os.path.join(dst, f), 'rb') as fh:
            manifest.append('%s  %s' % (hashlib.sha256(fh.read()).hexdigest(), f))
    with open(os.path.join(dst, '__init__.py'), 'w') as fh:      # what setup.py:31-43 generates
        fh.write("__abspath__ = %r\n" % (REF_DST.rstrip('/') + '/'))
    with open(os.path.join(REF_DST, 'MANIFEST.sha256'), 'w') as fh:
        fh.write('\n'.join(manifest) + '\n')
    if verbose:
        print('oracle/_ref: %d reference modules copied unmodified from %s' % (len(FILES), src))
    return REF_DST


if __name__ == '__main__':
    sys.exit(0 if make() else 1)
