"""Device-resident model + batched forward model / likelihood (thin layer over
the C-ABI).  PyTorch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Flags, GridDesc, PhotNetDesc, SpecNetDesc, check


def load_container(obj, kind=None, bands=None):
    """A model container is a dict of arrays or a path to an ``.npz``; an HDF5 path (what an unmodified caller of
    the reference passes: fastMISTmod.py:65-66, genmod.py:15-44) is converted on the fly where h5py is importable
    (minesweeper_b200.ingest; ``kind`` in 'mist' | 'phot' | 'spec')."""
    if obj is None:
        return None
    if isinstance(obj, dict):
        return obj
    import os
    name = os.fspath(obj)
    if kind is not None and (name.endswith(('.h5', '.hdf5')) or (kind == 'phot' and os.path.isdir(name))):
        from . import ingest
        return ingest.load_model_file(name, kind, bands=bands)
    with np.load(obj, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def _dp(a):
    return a.ctypes.data_as(_lib.c_dp)


def _fp(a):
    return a.ctypes.data_as(_lib.c_fp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Engine(object):
    """Owns one ``ms_ctx`` on one GPU.

    mist / photnet / specnet are containers as produced by
    ``minesweeper_b200.synth`` (or ``.npz`` paths); any may be None.
    ``bands`` selects and orders the photometric bands (default: all).
    """

    def __init__(self, mist=None, photnet=None, specnet=None, bands=None, precision='f32',
                 device=None):
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device_index = int(device)
        self.precision = precision
        keep = []
        gd = pd = sd = None
        mist = load_container(mist)
        if mist is not None:
            ax = [_f64(mist[k]) for k in ('eep', 'mass', 'feh', 'afe')]
            vals = _f64(mist['values'])
            valid = np.ascontiguousarray(mist['valid'], dtype=np.uint8)
            gd = GridDesc(len(ax[0]), len(ax[1]), len(ax[2]), len(ax[3]), _dp(ax[0]), _dp(ax[1]),
                          _dp(ax[2]), _dp(ax[3]), _dp(vals), valid.ctypes.data_as(_lib.c_u8p))
            keep += ax + [vals, valid]
            self.grid_shape = tuple(len(a) for a in ax)
        photnet = load_container(photnet)
        self.bands = []
        if photnet is not None:
            names = [str(b) for b in photnet['bands']]
            if bands is None:
                pick = list(range(len(names)))
            else:
                missing = [str(b) for b in bands if str(b) not in names]
                if missing:          # the scalar genphot would raise KeyError on them; do not drop them silently
                    raise KeyError('observed bands %s are not in the photometric net (%s)' % (missing, names))
                pick = [names.index(str(b)) for b in bands]
            self.bands = [names[i] for i in pick]
            w = {k: _f32(photnet[k][pick]) for k in ('w1', 'b1', 'w2', 'b2', 'w3', 'b3')}
            xmin, xmax = _f64(photnet['xmin']), _f64(photnet['xmax'])
            nb, h, d_in = w['w1'].shape
            pd = PhotNetDesc(nb, d_in, h, _fp(w['w1']), _fp(w['b1']), _fp(w['w2']), _fp(w['b2']),
                             _fp(w['w3']), _fp(w['b3']), _dp(xmin), _dp(xmax))
            keep += list(w.values()) + [xmin, xmax]
            self.phot_h = h
        specnet = load_container(specnet)
        if specnet is not None:
            w = {k: _f32(specnet[k]) for k in ('w0', 'b0', 'w1', 'b1', 'w2', 'b2')}
            xmin, xmax = _f64(specnet['xmin']), _f64(specnet['xmax'])
            wave = _f64(specnet['wavelength'])
            h, d_in = w['w0'].shape
            npix = w['w2'].shape[0]
            sd = SpecNetDesc(d_in, h, npix, _fp(w['w0']), _fp(w['b0']), _fp(w['w1']), _fp(w['b1']),
                             _fp(w['w2']), _fp(w['b2']), _dp(xmin), _dp(xmax), _dp(wave),
                             float(specnet['resolution']), float(specnet['leak']),
                             int(bool(specnet['logt_input'])))
            keep += list(w.values()) + [xmin, xmax, wave]
            self.spec_shape = (d_in, h, npix)
            self.spec_wave = wave.copy()
        self._ctx = C.c_void_p()
        check(self.lib.ms_create(C.byref(self._ctx), self.device_index, _lib.PRECISIONS[precision],
                                 C.byref(gd) if gd else None, C.byref(pd) if pd else None,
                                 C.byref(sd) if sd else None))
        del keep
        self.device = torch.device('cuda', self.device_index)
        self.n_obs = {}
        self.nslots_phot = 0            # star slots with photometry loaded: 0 .. nslots_phot - 1

    def close(self):
        if getattr(self, '_ctx', None) is not None and self._ctx.value:
            self.lib.ms_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- observations ------------------------------------------------------
    def set_star(self, slot=0, obs_phot=None, obs_wave=None, obs_flux=None, obs_eflux=None):
        """obs_phot: ``{band: [mag, err]}`` (reference fitargs['obs_phot']) or a
        pair of arrays in ``self.bands`` order; spectrum arrays as
        fitargs['obs_wave_fit' | 'obs_flux_fit' | 'obs_eflux_fit']."""
        nb = 0
        mag = err = None
        if obs_phot is not None:
            if isinstance(obs_phot, dict):
                mag = np.full(len(self.bands), np.nan)
                err = np.full(len(self.bands), np.nan)
                for i, b in enumerate(self.bands):
                    if b in obs_phot:
                        mag[i], err[i] = obs_phot[b][0], obs_phot[b][1]
            else:
                mag, err = _f64(obs_phot[0]), _f64(obs_phot[1])
            nb = len(mag)
        n_obs = 0
        w = f = e = None
        if obs_wave is not None:
            w, f, e = _f64(obs_wave), _f64(obs_flux), _f64(obs_eflux)
            n_obs = len(w)
        check(self.lib.ms_set_star(self._ctx, int(slot), _dp(mag) if nb else None,
                                   _dp(err) if nb else None, nb, _dp(w) if n_obs else None,
                                   _dp(f) if n_obs else None, _dp(e) if n_obs else None, n_obs))
        self.n_obs[int(slot)] = n_obs
        if nb:
            self.nslots_phot = max(self.nslots_phot, int(slot) + 1)

    def set_lsf(self, slot, sigma=None):
        """Wavelength-dependent LSF of a star (array-valued Inst_R, genmod.py:74-77): ``sigma[npix]`` on the spectral
        net's wavelength grid, in wavelength units; None clears.  See minesweeper_b200/lsf.py."""
        if sigma is None:
            check(self.lib.ms_set_lsf(self._ctx, int(slot), 0, None, None, None, 0.0))
            return
        from . import lsf
        nx, lam, idx, frac, xps = lsf.prepare(self.spec_wave, sigma)
        check(self.lib.ms_set_lsf(self._ctx, int(slot), nx, _dp(lam), idx.ctypes.data_as(_lib.c_ip), _dp(frac), xps))

    def set_stars_phot(self, obs_mag, obs_err, first_slot=0):
        """Catalog mode: obs_mag / obs_err[nstars, nb] (NaN = unobserved) into consecutive slots."""
        m, e = _f64(obs_mag), _f64(obs_err)
        check(self.lib.ms_set_stars_phot(self._ctx, int(first_slot), m.shape[0], _dp(m), _dp(e), m.shape[1]))
        self.nslots_phot = max(self.nslots_phot, int(first_slot) + m.shape[0])

    def check_star_slots(self, nstars, star_plx=None):
        """Catalog callers (the samplers) address star slots 0 .. nstars - 1: they must all be loaded, and a
        per-star parallax table must cover them (the kernels index both by slot)."""
        if nstars > 1 and nstars > self.nslots_phot:
            raise ValueError('%d stars requested but only %d star slots have photometry loaded '
                             '(set_stars_phot / set_star)' % (nstars, self.nslots_phot))
        if star_plx is not None and star_plx.shape[0] < nstars:
            raise ValueError('star_plx has %d rows for %d stars' % (star_plx.shape[0], nstars))

    def reserve(self, max_batch, lock=True):
        """Pre-size the ctx's per-batch scratch for batches of up to ``max_batch`` rows and (lock=True) pin it, so
        that CUDA graphs capturing lnlike_batch / lnprob_batch keep valid pointers; ``reserve(0, lock=False)`` lifts
        the pin.  See ms_reserve in include/minesweeper_b200.h."""
        check(self.lib.ms_reserve(self._ctx, int(max_batch), int(bool(lock))))

    # -- helpers -------------------------------------------------------------
    def _dev(self, a):
        t = torch.as_tensor(a, dtype=torch.float64)
        return t.to(self.device, non_blocking=True).contiguous()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- forward-model pieces ----------------------------------------------
    def mist_interp(self, theta4):
        """theta4[B,4] = (EEP, initial_Mass, initial_[Fe/H], initial_[a/Fe]).
        Returns (brackets[B,4] int32, derived[B,13] f64, ok[B] uint8) on the GPU."""
        th = self._dev(theta4)
        B = th.shape[0]
        br = torch.empty((B, 4), dtype=torch.int32, device=self.device)
        der = torch.empty((B, _lib.MS_NDERIVED), dtype=torch.float64, device=self.device)
        ok = torch.empty((B,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ms_mist_interp(self._ctx, th.data_ptr(), B, br.data_ptr(), der.data_ptr(),
                                          ok.data_ptr(), self._stream()))
        return br, der, ok

    def phot_sed(self, photpars):
        """photpars[B,7] = (Teff, logg, feh, afe, logR, Dist, Av) -> mags[B,nb]."""
        pp = self._dev(photpars)
        B = pp.shape[0]
        mags = torch.empty((B, len(self.bands)), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ms_phot_sed(self._ctx, pp.data_ptr(), B, mags.data_ptr(), self._stream()))
        return mags

    def spec_model(self, specpars, slot=0):
        """specpars[B, 8 + n_pc] -> model flux[B, n_obs] on the slot's obs_wave."""
        sp = self._dev(specpars)
        B, nc = sp.shape
        flux = torch.empty((B, self.n_obs[slot]), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ms_spec_model(self._ctx, int(slot), sp.data_ptr(), nc - 8, B,
                                         flux.data_ptr(), self._stream()))
        return flux

    # -- the likelihood -----------------------------------------------------
    @staticmethod
    def make_colmap(fitpars_i):
        return np.array([_lib.param_id(p) for p in fitpars_i], dtype=np.int32)

    @staticmethod
    def make_fixed(fixedpars):
        fx = np.full(_lib.MS_NPARAM, np.nan)
        for k, v in (fixedpars or {}).items():
            try:
                fx[_lib.param_id(k)] = float(v)
            except KeyError:
                pass                       # fixed parameters the forward model never reads
        return fx

    @staticmethod
    def make_flags(spec=False, phot=True, modpoly=False, ageweight=True, mistdif=True, n_pc=0, slot=0):
        return Flags(int(spec), int(phot), int(modpoly), int(ageweight), int(mistdif), int(n_pc), int(slot))

    def lnlike_batch(self, theta, colmap, fixed, flags, star_slot=None, want_derived=True,
                     want_brackets=False, out=None, derived_out=None):
        """theta[B,ndim] (GPU f64 tensor or array) -> (lnL[B], derived[B,13] | None, brackets | None)."""
        th = self._dev(theta)
        B, ndim = th.shape
        lnl = out if out is not None else torch.empty((B,), dtype=torch.float64, device=self.device)
        iso = _lib.PARAM_IDS['EEP'] in list(colmap) or not np.isnan(fixed[_lib.PARAM_IDS['EEP']])
        der = derived_out if derived_out is not None else (
            torch.empty((B, _lib.MS_NDERIVED), dtype=torch.float64, device=self.device)
            if (want_derived and iso) else None)
        br = torch.empty((B, 4), dtype=torch.int32, device=self.device) if (want_brackets and iso) else None
        slots = None
        if star_slot is not None:
            slots = torch.as_tensor(star_slot, dtype=torch.int32).to(self.device).contiguous()
        cm = np.ascontiguousarray(colmap, dtype=np.int32)
        fx = _f64(fixed)
        with torch.cuda.device(self.device):
            check(self.lib.ms_lnlike_batch(
                self._ctx, C.byref(flags), th.data_ptr(), slots.data_ptr() if slots is not None else None,
                B, ndim, cm.ctypes.data_as(_lib.c_ip), _dp(fx), lnl.data_ptr(),
                der.data_ptr() if der is not None else None,
                br.data_ptr() if br is not None else None, self._stream()))
        return lnl, der, br

    def lnlike_batch_host(self, theta, colmap, fixed, flags, want_derived=False, lnl_out=None,
                          derived_out=None):
        """Host (NumPy, ideally pinned) in, host out; copies are inside the call."""
        th = _f64(theta)
        B, ndim = th.shape
        lnl = lnl_out if lnl_out is not None else np.empty(B)
        der = derived_out if derived_out is not None else \
            (np.empty((B, _lib.MS_NDERIVED)) if want_derived else None)
        cm = np.ascontiguousarray(colmap, dtype=np.int32)
        fx = _f64(fixed)
        check(self.lib.ms_lnlike_batch_host(self._ctx, C.byref(flags), _dp(th), B, ndim,
                                            cm.ctypes.data_as(_lib.c_ip), _dp(fx), _dp(lnl),
                                            _dp(der) if der is not None else None))
        return lnl, der

    def launch_count(self):
        return int(self.lib.ms_launch_count(self._ctx))

    def set_fusion(self, on=True):
        """Tensor-core modes: run photometry-only isochrone batches as one kernel (default) or as the
        interpolation kernel followed by the MLP kernel."""
        check(self.lib.ms_set_fusion(self._ctx, int(bool(on))))

    def set_lnl_peers(self, ptrs=(), offset=0):
        """Fused gather: the tensor-core photometry kernel also stores lnL[p] at ``ptrs[r] + 8 * (offset + p)`` for every
        rank r (device pointers to peer-mapped buffers, see distributed.PeerGather); ``set_lnl_peers()`` clears."""
        arr = (C.c_uint64 * max(len(ptrs), 1))(*[int(p) for p in ptrs])
        check(self.lib.ms_set_lnl_peers(self._ctx, arr, len(ptrs), int(offset)))

    def set_spec_fast(self, on=True):
        """f32-class modes: fast broadening kernel (default) or the generic FFT kernel for every row."""
        check(self.lib.ms_set_spec_fast(self._ctx, int(on)))

    def profile(self, on=True):
        check(self.lib.ms_profile_enable(self._ctx, int(on)))

    def profile_read(self):
        """{kernel family: (device ms summed, launches)} since profile(True)."""
        n = len(_lib.KERNEL_KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        check(self.lib.ms_profile_read(self._ctx, ms, cnt))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(_lib.KERNEL_KINDS)}
