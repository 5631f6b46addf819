"""Host-side mirror of the reference's ``GenMIST`` (reference
minesweeper/fastMISTmod.py:26-108) on top of the CUDA interpolator.

Same constructor keywords (``MISTpath``, ``ageweight``), same attributes a caller
reads (``modpararr``, ``gridpoints``, ``binwidths``, ``labels``, ``predictions``)
and the same ``getMIST(mass=, eep=, feh=, afe=)`` -> list of 13 | ``None``; plus
``getMIST_batch`` for arrays.  ``MISTpath`` is a dense-grid container
(``minesweeper_b200.synth.make_mist`` layout, dict or ``.npz``), see
INTEGRATION.md for converting the reference's HDF5 track tables.
"""
import numpy as np

from .engine import Engine, load_container

LABELS = ['EEP', 'initial_Mass', 'initial_[Fe/H]', 'initial_[a/Fe]']
PREDICTIONS = ['log(Age)', 'Mass', 'log(R)', 'log(L)', 'log(Teff)', '[Fe/H]', '[a/Fe]', 'log(g)']


class GenMIST(object):
    def __init__(self, **kwargs):
        self.verbose = kwargs.get('verbose', True)
        self.mistfile = kwargs.get('MISTpath', None)
        if self.mistfile is None:
            raise IOError('MISTpath is required: the MIST tables are not shipped with the package')
        self.ageweight = kwargs.get('ageweight', True)
        self.labels = list(LABELS)
        self.predictions = list(PREDICTIONS) + (['Agewgt'] if self.ageweight else [])
        self.ndim = len(self.labels)
        self.modpararr = self.labels + self.predictions
        grid = load_container(self.mistfile)
        self.gridpoints = {p: np.asarray(grid[k], dtype=np.float64)
                           for p, k in zip(self.labels, ('eep', 'mass', 'feh', 'afe'))}
        self.binwidths = {p: np.diff(v) for p, v in self.gridpoints.items()}
        self.engine = kwargs.get('engine', None)
        if self.engine is None:
            self.engine = Engine(mist=grid, precision=kwargs.get('precision', 'f32'),
                                 device=kwargs.get('device', None))
        self._ncol = len(self.modpararr)

    def getMIST_batch(self, eep, mass, feh, afe):
        """Arrays in -> (pred[B, 12|13] f64 NumPy in ``modpararr`` order, ok[B] bool,
        brackets[B,4] int32).  Rows with ``ok=False`` are the reference's ``None``."""
        th = np.stack([np.asarray(v, dtype=np.float64).ravel() for v in (eep, mass, feh, afe)], axis=1)
        br, der, ok = self.engine.mist_interp(th)
        return der.cpu().numpy()[:, :self._ncol], ok.cpu().numpy().astype(bool), br.cpu().numpy()

    def getMIST(self, mass=1.0, eep=300, feh=0.0, afe=0.0, **kwargs):
        pred, ok, _ = self.getMIST_batch([eep], [mass], [feh], [afe])
        if not ok[0]:
            return None
        return [eep, mass, feh, afe] + list(pred[0, 4:])
