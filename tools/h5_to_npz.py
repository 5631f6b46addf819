#!/usr/bin/env python
"""Offline converter: the reference's HDF5 model files -> the .npz containers minesweeper_b200 loads.
Run where h5py is installed (it is not a dependency of the package, and is absent from the build image).

    python tools/h5_to_npz.py mist  MIST_2.0_EEPtrk.h5            mist.npz
    python tools/h5_to_npz.py spec  YSTANN.h5                     specnet.npz  [--leak 0.01] [--logt]
    python tools/h5_to_npz.py phot  /path/to/nnMIST/              photnet.npz  --bands 2MASS_J 2MASS_H ...

Layouts: MIST track tables as GenMIST.make_lib reads them (reference minesweeper/fastMISTmod.py:65-98); Payne network
files as recalled in SURVEY.md Appendix B (lin1/lin2/lin3 per filter; w_array_*/b_array_*/x_min/x_max/wavelength/
resolution), every shape taken from the file.  See minesweeper_b200/ingest.py."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('kind', choices=['mist', 'spec', 'phot'])
    ap.add_argument('src')
    ap.add_argument('dst')
    ap.add_argument('--bands', nargs='*', default=None)
    ap.add_argument('--leak', type=float, default=0.01)
    ap.add_argument('--logt', action='store_true', help='first spectral label is log10(Teff)')
    ap.add_argument('--no-ageweight', action='store_true')
    a = ap.parse_args()
    from minesweeper_b200 import ingest
    if a.kind == 'mist':
        c = ingest.mist_dense_from_h5(a.src, ageweight=not a.no_ageweight)
    elif a.kind == 'spec':
        c = ingest.load_model_file(a.src, 'spec', leak=a.leak, logt_input=a.logt)
    else:
        bands = a.bands or [f[len('nnMIST_'):-3] for f in sorted(os.listdir(a.src)) if f.startswith('nnMIST_') and f.endswith('.h5')]
        c = ingest.load_model_file(a.src, 'phot', bands=bands)
    np.savez(a.dst, **c)
    print('%s: wrote %s (%s)' % (a.kind, a.dst, ', '.join('%s%s' % (k, getattr(v, 'shape', '')) for k, v in c.items())))


if __name__ == '__main__':
    main()
