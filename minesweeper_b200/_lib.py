"""ctypes binding of libmsb200.so (include/minesweeper_b200.h).

There is no CPU fallback: if the library is missing this module raises, and
``ms_create`` itself fails with MS_ENODEV when no sm_100 device is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MS_LIB_PATH') or os.path.join(HERE, 'libmsb200.so')   # MS_LIB_PATH: tuning variants only

MS_OK, MS_EINVAL, MS_ENODEV, MS_ECUDA, MS_ENOMEM, MS_ESTATE = 0, -1, -2, -3, -4, -5
MS_PREC_F64, MS_PREC_F32, MS_PREC_TC, MS_PREC_TC16 = 0, 1, 2, 3
PRECISIONS = {'f64': MS_PREC_F64, 'f32': MS_PREC_F32, 'tc': MS_PREC_TC, 'tc16': MS_PREC_TC16}

# canonical parameter ids (fitstar.py:50-67 order)
PARAM_IDS = {
    'Teff': 0, 'log(g)': 1, '[Fe/H]': 2, '[a/Fe]': 3, 'Vrad': 4, 'Vrot': 5, 'Vmic': 6,
    'Inst_R': 7, 'log(R)': 8, 'Dist': 9, 'Av': 10, 'log(A)': 11, 'EEP': 12,
    'initial_[Fe/H]': 13, 'initial_[a/Fe]': 14, 'initial_Mass': 15,
}
MS_P_PC0 = 16
MS_MAX_PC = 8
MS_NPARAM = MS_P_PC0 + MS_MAX_PC
MS_NDERIVED = 13
DERIVED_NAMES = ('EEP', 'initial_Mass', 'initial_[Fe/H]', 'initial_[a/Fe]', 'log(Age)', 'Mass',
                 'log(R)', 'log(L)', 'log(Teff)', '[Fe/H]', '[a/Fe]', 'log(g)', 'Agewgt')


def param_id(name):
    if name in PARAM_IDS:
        return PARAM_IDS[name]
    if name.startswith('pc_'):
        i = int(name[3:])
        if 0 <= i < MS_MAX_PC:
            return MS_P_PC0 + i
    raise KeyError('parameter %r is not known to the batched likelihood' % name)


c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)


class GridDesc(C.Structure):
    _fields_ = [('ne', C.c_int32), ('nm', C.c_int32), ('nf', C.c_int32), ('na', C.c_int32),
                ('eep', c_dp), ('mass', c_dp), ('feh', c_dp), ('afe', c_dp),
                ('values', c_dp), ('valid', c_u8p)]


class PhotNetDesc(C.Structure):
    _fields_ = [('nb', C.c_int32), ('d_in', C.c_int32), ('h', C.c_int32),
                ('w1', c_fp), ('b1', c_fp), ('w2', c_fp), ('b2', c_fp), ('w3', c_fp), ('b3', c_fp),
                ('xmin', c_dp), ('xmax', c_dp)]


class SpecNetDesc(C.Structure):
    _fields_ = [('d_in', C.c_int32), ('h', C.c_int32), ('npix', C.c_int32),
                ('w0', c_fp), ('b0', c_fp), ('w1', c_fp), ('b1', c_fp), ('w2', c_fp), ('b2', c_fp),
                ('xmin', c_dp), ('xmax', c_dp), ('wavelength', c_dp),
                ('resolution', C.c_double), ('leak', C.c_double), ('logt_input', C.c_int32)]


class Flags(C.Structure):
    _fields_ = [('spec', C.c_int32), ('phot', C.c_int32), ('modpoly', C.c_int32),
                ('ageweight', C.c_int32), ('mistdif', C.c_int32), ('n_pc', C.c_int32), ('slot', C.c_int32)]


class PriorCol(C.Structure):
    _fields_ = [('kind', C.c_int32), ('pad_', C.c_int32), ('p', C.c_double * 4)]


class PriorTerm(C.Structure):
    _fields_ = [('value_id', C.c_int32), ('kind', C.c_int32), ('p', C.c_double * 4)]


class AdvPrior(C.Structure):
    _fields_ = [('gal', C.c_int32), ('galage', C.c_int32), ('vrot', C.c_int32), ('alpha', C.c_int32),
                ('Xp', C.c_double), ('Yp', C.c_double), ('Zp', C.c_double), ('sol_X', C.c_double), ('sol_Z', C.c_double),
                ('age_thin', C.c_double * 4), ('age_thick', C.c_double * 4), ('age_halo', C.c_double * 4),
                ('vrot_dwarf', C.c_double * 3), ('vrot_giant', C.c_double * 3), ('minalpha', C.c_double)]


PRIOR_KINDS = {'uniform': 0, 'gaussian': 1, 'tgaussian': 2, 'exp': 3, 'loguniform': 4, 'beta': 5, 'table': 6}
TERM_KINDS = {'uniform': 0, 'gaussian': 1, 'tgaussian': 2, 'tgaussian_bounds': 3}
MS_V_LOGAGE, MS_V_AGE, MS_V_PARALLAX = MS_NPARAM, MS_NPARAM + 1, MS_NPARAM + 2
MS_MAX_PRIOR_TERMS = 32
MS_NS_TRAIL = 7                     # trailing columns of a dead-point record (ms_ns_state.dead)

class NsState(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ('S', 'K', 'Q', 'nd', 'nder', 'walks')] +
                [('maxiter', C.c_int64), ('dead_cap', C.c_int64), ('dlogz', C.c_double)] +
                [(n, C.c_void_p) for n in ('live_u', 'live_lp', 'live_row', 'cur_u', 'cur_lp', 'cur_row', 'prop_u',
                                           'prop_lp', 'prop_theta', 'prop_der', 'nacc', 'lmin', 'sigma', 'scale',
                                           'logz', 'hacc', 'wmax', 'wtot', 'wsum', 'wsq', 'niter', 'ncall', 'done', 'dead',
                                           'rng_round', 'center', 'ell_r')] +
                [('unif', C.c_int32)])


# every exported symbol of include/minesweeper_b200.h with its prototype
PROTOTYPES = {
    'ms_version': (C.c_char_p, []),
    'ms_last_error': (C.c_char_p, []),
    'ms_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(GridDesc),
                            C.POINTER(PhotNetDesc), C.POINTER(SpecNetDesc)]),
    'ms_destroy': (None, [C.c_void_p]),
    'ms_precision': (C.c_int, [C.c_void_p]),
    'ms_set_star': (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp, C.c_int]),
    'ms_mist_interp': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    'ms_phot_sed': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'ms_spec_model': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                C.c_void_p]),
    'ms_lnlike_batch': (C.c_int, [C.c_void_p, C.POINTER(Flags), C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_int, c_ip, c_dp, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    'ms_lnlike_batch_host': (C.c_int, [C.c_void_p, C.POINTER(Flags), c_dp, C.c_int64, C.c_int, c_ip,
                                       c_dp, c_dp, c_dp]),
    'ms_prior_transform': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(PriorCol),
                                     C.c_void_p, C.c_void_p]),
    'ms_lnprob_batch': (C.c_int, [C.c_void_p, C.POINTER(Flags), C.c_void_p, C.c_int64, C.c_int, c_ip, c_dp,
                                  C.c_void_p, C.POINTER(PriorTerm), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    'ms_lnprob_batch_stars': (C.c_int, [C.c_void_p, C.POINTER(Flags), C.c_void_p, C.c_int64, C.c_int, c_ip, c_dp,
                                        C.c_void_p, C.POINTER(PriorTerm), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    'ms_set_lsf': (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_ip, c_dp, C.c_double]),
    'ms_set_prior_table': (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int]),
    'ms_lnprob_batch_adv': (C.c_int, [C.c_void_p, C.POINTER(Flags), C.c_void_p, C.c_int64, C.c_int, c_ip, c_dp,
                                      C.c_void_p, C.POINTER(PriorTerm), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(AdvPrior), C.c_void_p]),
    'ms_set_stars_phot': (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, C.c_int]),
    'ms_ns_propose': (C.c_int, [C.c_void_p, C.POINTER(NsState), C.c_uint64, C.c_uint64, C.c_void_p]),
    'ms_ns_accept': (C.c_int, [C.c_void_p, C.POINTER(NsState), C.c_void_p]),
    'ms_ns_propose_transform': (C.c_int, [C.c_void_p, C.POINTER(NsState), C.c_uint64, C.c_uint64, C.POINTER(PriorCol),
                                          C.c_void_p]),
    'ms_ns_lnprob_accept': (C.c_int, [C.c_void_p, C.POINTER(NsState), C.POINTER(Flags), c_ip, c_dp,
                                      C.POINTER(PriorTerm), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(AdvPrior), C.c_void_p]),
    'ms_ns_consume': (C.c_int, [C.c_void_p, C.POINTER(NsState), C.c_uint64, C.c_void_p]),
    'ms_ns_finalize': (C.c_int, [C.c_void_p, C.POINTER(NsState), C.c_void_p]),
    'ms_ns_spread': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    'ms_reserve': (C.c_int, [C.c_void_p, C.c_int64, C.c_int]),
    'ms_launch_count': (C.c_int64, [C.c_void_p]),
    'ms_set_fusion': (C.c_int, [C.c_void_p, C.c_int]),
    'ms_set_lnl_peers': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int64]),
    'ms_set_spec_fast': (C.c_int, [C.c_void_p, C.c_int]),
    'ms_profile_enable': (C.c_int, [C.c_void_p, C.c_int]),
    'ms_profile_read': (C.c_int, [C.c_void_p, c_dp, C.POINTER(C.c_int64)]),
}
KERNEL_KINDS = ('interp', 'pars', 'phot', 'spec_mlp', 'spec_smooth', 'finalize')

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'minesweeper_b200: %s not found. Build it with `python -m minesweeper_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MSError(RuntimeError):
    pass


def check(rc):
    if rc != MS_OK:
        msg = load().ms_last_error()
        raise MSError('libmsb200 error %d: %s' % (rc, msg.decode() if msg else ''))
