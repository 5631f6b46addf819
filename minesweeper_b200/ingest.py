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


# ---- Payne network weight files (genmod.py:15-44 hands their paths to the un-vendored Payne package) --------------
# Layouts as recalled in SURVEY.md Appendix B (PARITY UNPINNED, like the ANN arithmetic itself); every shape is read
# from the file, nothing is assumed about widths.  Inputs are "HDF5-like" mappings (an open h5py.File / Group, or nested
# dicts of arrays in tests); `open_h5` turns a path into one where h5py is importable.
def open_h5(path):
    try:
        import h5py
    except ImportError as e:
        raise ImportError('reading %r needs h5py, which this package does not depend on: run tools/h5_to_npz.py on a '
                          'machine that has it and pass the .npz' % (path,)) from e
    return h5py.File(path, 'r')                                                      # pragma: no cover


def _flatten(group, prefix=''):
    """{'a/b/c': array} for every dataset under an HDF5-like mapping."""
    out = {}
    for k in group.keys():
        v = group[k]
        name = prefix + _name(k)
        if hasattr(v, 'keys') and not isinstance(v, np.ndarray):
            out.update(_flatten(v, name + '/'))
        else:
            out[name] = np.asarray(v)
    return out


def _find(flat, *needles):
    hits = [k for k in flat if all(n in k for n in needles)]
    if len(hits) != 1:
        raise KeyError('expected exactly one dataset matching %s, found %s among %s' % (needles, hits, sorted(flat)))
    return flat[hits[0]]


def payne_phot_net(files, bands=None):
    """Per-filter photometric nets -> the stacked container of ``synth.make_phot_net`` / ms_phot_net_desc.

    ``files``: ``{band: HDF5-like mapping}`` of the ``nnMIST_<band>.h5`` files (reference data/nnMIST/README.md;
    FastPayneSEDPredict(usebands, nnpath), genmod.py:39-44), each holding a 3-layer perceptron
    ``lin1 / lin2 / lin3`` (``weight`` [out, in], ``bias``) and the input box ``xmin`` / ``xmax``."""
    bands = list(files.keys()) if bands is None else [b for b in bands if b in files]
    if not bands:
        raise ValueError('no photometric net files for the requested bands')
    per = []
    for b in bands:
        flat = _flatten(files[b])
        per.append(dict(w1=_find(flat, 'lin1', 'weight'), b1=_find(flat, 'lin1', 'bias'),
                        w2=_find(flat, 'lin2', 'weight'), b2=_find(flat, 'lin2', 'bias'),
                        w3=_find(flat, 'lin3', 'weight'), b3=_find(flat, 'lin3', 'bias'),
                        xmin=_find(flat, 'xmin'), xmax=_find(flat, 'xmax')))
    h, d_in = per[0]['w1'].shape
    for b, p in zip(bands, per):
        if p['w1'].shape != (h, d_in) or p['w2'].shape != (h, h) or p['w3'].shape != (1, h):
            raise ValueError('net of band %s has a different architecture (%s)' % (b, p['w1'].shape))
        if not (np.allclose(p['xmin'], per[0]['xmin']) and np.allclose(p['xmax'], per[0]['xmax'])):
            raise ValueError('net of band %s was trained on a different input box' % b)
    f32 = np.float32
    return dict(bands=np.array(bands), w1=np.stack([p['w1'] for p in per]).astype(f32),
                b1=np.stack([p['b1'] for p in per]).astype(f32), w2=np.stack([p['w2'] for p in per]).astype(f32),
                b2=np.stack([p['b2'] for p in per]).astype(f32), w3=np.stack([p['w3'] for p in per]).astype(f32),
                b3=np.stack([p['b3'].reshape(1) for p in per]).astype(f32),
                xmin=np.asarray(per[0]['xmin'], dtype=np.float64).ravel()[:d_in],
                xmax=np.asarray(per[0]['xmax'], dtype=np.float64).ravel()[:d_in], activation=np.array('sigmoid'))


def payne_spec_net(h5like, leak=0.01, logt_input=False):
    """Spectral net file (``w_array_0/1/2``, ``b_array_0/1/2``, ``x_min``, ``x_max``, ``wavelength``, ``resolution``;
    PayneSpecPredict(nnpath), genmod.py:24-34) -> the container of ``synth.make_spec_net`` / ms_spec_net_desc."""
    flat = _flatten(h5like)
    w = [_find(flat, 'w_array_%d' % i) for i in range(3)]
    b = [_find(flat, 'b_array_%d' % i) for i in range(3)]
    d_in = w[0].shape[1]
    if w[1].shape != (w[0].shape[0],) * 2 or w[2].shape[1] != w[0].shape[0]:
        raise ValueError('unexpected spectral net shapes %s' % ([x.shape for x in w],))
    wave = np.asarray(_find(flat, 'wavelength'), dtype=np.float64).ravel()
    if len(wave) != w[2].shape[0]:
        raise ValueError('wavelength grid (%d) does not match the output layer (%d)' % (len(wave), w[2].shape[0]))
    f32 = np.float32
    return dict(w0=w[0].astype(f32), b0=b[0].astype(f32).ravel(), w1=w[1].astype(f32), b1=b[1].astype(f32).ravel(),
                w2=w[2].astype(f32), b2=b[2].astype(f32).ravel(),
                xmin=np.asarray(_find(flat, 'x_min'), dtype=np.float64).ravel()[:d_in],
                xmax=np.asarray(_find(flat, 'x_max'), dtype=np.float64).ravel()[:d_in], wavelength=wave,
                resolution=np.float64(np.asarray(_find(flat, 'resolution')).ravel()[0]), leak=np.float64(leak),
                logt_input=np.bool_(logt_input))


def load_model_file(path, kind, bands=None, **kw):
    """``fitargs['MISTpath' | 'photANNpath' | 'specANNpath']`` given as an HDF5 path, the way an unmodified caller of
    the reference passes them (fastMISTmod.py:65-66, genmod.py:15-44): read with h5py where it is importable.
    kind: 'mist' | 'phot' | 'spec'.  For 'phot', ``path`` is the directory (or prefix) holding ``nnMIST_<band>.h5``."""
    import os
    if kind == 'mist':
        return mist_dense_from_h5(path, **kw)
    if kind == 'spec':
        with open_h5(path) as f:                                                     # pragma: no cover
            return payne_spec_net(f, **kw)
    if kind == 'phot':                                                               # pragma: no cover
        files = {}
        for b in bands or []:
            p = os.path.join(path, 'nnMIST_%s.h5' % b)
            if os.path.exists(p):
                files[b] = open_h5(p)
        try:
            return payne_phot_net(files, bands)
        finally:
            for f in files.values():
                f.close()
    raise ValueError('unknown model kind %r' % kind)
