"""Model-file ingestion: MIST track tables -> the dense grid container the engine loads.

Restates what ``GenMIST.make_lib`` / ``lib_as_grid`` do to the HDF5 input (reference
minesweeper/fastMISTmod.py:76-98, 110-125): concatenate the groups listed under ``'index'``,
round mass / [Fe/H] / [a/Fe] to 2 decimals, take the unique label values as grid axes and place
every table row at its integer grid coordinate.  The result is ``values[na, nf, nm, ne, 9]`` in
``GenMIST.output`` column order plus ``valid[na, nf, nm, ne]``, i.e. exactly
what the interpolation kernel gathers from; vertices absent from the tables stay invalid and are
handled by the weight renormalisation (fastMISTmod.py:164-169).

The input is any mapping with the HDF5 file's structure: ``tables['index']`` is the list of group
names and ``tables[name]`` a structured array with the MIST column names.  An open ``h5py.File``
satisfies that, and so does ``minesweeper_b200.synth.mist_rows`` (used by the tests; h5py is not a
dependency of this package).
"""
import numpy as np

LABEL_FIELDS = ('EEP', 'initial_mass', 'initial_[Fe/H]', 'initial_[a/Fe]')           # fastMISTmod.py:80
OUTPUT_FIELDS = ('log_age', 'star_mass', 'log_R', 'log_L', 'log_Teff', '[Fe/H]', '[a/Fe]', 'log_g',
                 'Agewgt')                                                           # :84-87, 57-60


def _name(z):
    return z.decode() if isinstance(z, bytes) else str(z)


def mist_dense_from_tables(tables, ageweight=True):
    """HDF5-like MIST tables -> ``dict(eep, mass, feh, afe, values, valid)``.

    ``ageweight=False`` fills the Agewgt column with ones (the reference then does not read it,
    fastMISTmod.py:57-60).  A grid coordinate that occurs in more than one row keeps the last row
    (the reference would average such duplicates through the weights; real MIST tables have none).
    """
    groups = [np.asarray(tables[_name(z)]) for z in tables['index']]
    if not groups:
        raise ValueError('MIST tables: empty index')
    lab = [np.concatenate([g[f] for g in groups]).astype(np.float64) for f in LABEL_FIELDS]
    lab[1], lab[2], lab[3] = (np.around(x, decimals=2) for x in lab[1:])             # :89-91
    cols = []
    for f in OUTPUT_FIELDS:
        if f == 'Agewgt' and not ageweight:
            cols.append(np.ones(len(lab[0])))
        else:
            cols.append(np.concatenate([g[f] for g in groups]).astype(np.float64))
    out = np.stack(cols, 1)
    axes = [np.unique(x) for x in lab]                                               # :118
    idx = [np.searchsorted(a, x) for a, x in zip(axes, lab)]                         # exact members
    ne, nm, nf, na = (len(a) for a in axes)
    values = np.zeros((na, nf, nm, ne, len(OUTPUT_FIELDS)))
    valid = np.zeros((na, nf, nm, ne), dtype=bool)
    values[idx[3], idx[2], idx[1], idx[0]] = out
    valid[idx[3], idx[2], idx[1], idx[0]] = True
    return dict(eep=axes[0], mass=axes[1], feh=axes[2], afe=axes[3], values=values, valid=valid)


def mist_dense_from_h5(path, ageweight=True, save=None):
    """Read a MIST HDF5 file (needs h5py, which this package does not depend on) and optionally
    write the dense container as ``.npz`` for ``fitargs['MISTpath']``."""
    try:
        import h5py
    except ImportError as e:                                                         # pragma: no cover
        raise ImportError('mist_dense_from_h5 needs h5py; convert on a machine that has it and pass '
                          'the .npz as MISTpath') from e
    with h5py.File(path, 'r') as f:                                                  # pragma: no cover
        tables = {'index': [_name(z) for z in f['index']]}
        for z in tables['index']:
            tables[z] = np.array(f[z])
    dense = mist_dense_from_tables(tables, ageweight=ageweight)                      # pragma: no cover
    if save:
        np.savez(save, **dense)
    return dense
