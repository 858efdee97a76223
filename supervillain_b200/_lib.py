"""ctypes binding of libsvb200.so -- the only way the Python layer reaches the GPU kernels.

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SVB200_LIB') or os.path.join(HERE, 'libsvb200.so')   # override: kernel tuning experiments

# mirrors include/svb200.h
F64, F32, I32, I64 = 0, 1, 2, 3
RNG_PHILOX, RNG_INJECTED = 0, 1
ARITH_STRICT, ARITH_FAST = 0, 1
PATH_AUTO, PATH_SMEM, PATH_GLOBAL, PATH_TILED = 0, 1, 2, 3
VOBS_ACTION, VOBS_SUM_DN2, VOBS_WRAP0, VOBS_WRAP1, VOBS_ACCEPTED, VOBS_ACCEPTANCE, VOBS_COUNT = range(7)
(WOBS_SUM_F2, WOBS_SUM_DF2, WOBS_WRAP0, WOBS_WRAP1, WOBS_ACCEPTED, WOBS_ACCEPTANCE, WOBS_DELTA_M_ABS,
 WOBS_COUNT) = range(8)
WL_JOINT, WL_VORTEX, WL_COEXACT = 0, 1, 2
OP_D, OP_DELTA, OP_FACE_SUM, OP_COFACE_SUM = 0, 1, 2, 3
CORR_SPIN, CORR_WINDING, CORR_VORTEX = 0, 1, 2
OVERLAP_PREDECESSOR = 1
VU_SITE, VU_LINK, VU_EXACT = 0, 1, 2

E_NULL, E_SHAPE, E_DTYPE, E_PARAM, E_UNSUPPORTED, E_ALIGN = -1, -2, -3, -4, -5, -6

# every symbol include/svb200.h declares, with its ctypes signature
_vp, _i, _i64, _u64, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
SIGNATURES = {
    'svb_version': (_i, []),
    'svb_last_error': (ctypes.c_char_p, []),
    'svb_villain_sweep': (_i, [_vp, _i, _vp, _i64, _i, _d, _vp, _i, _d, _i, _i, _u64, _u64, _u64, _i, _i, _i,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'svb_villain_sweep_overlapped': (_i, [_vp, _vp, _i64, _i, _d, _vp, _i, _d, _i, _i, _u64, _u64, _u64, _vp, _vp, _vp,
                                          ctypes.c_uint32, ctypes.c_uint32, _i, _vp]),
    'svb_villain_decoupled': (_i, [_i, _vp, _vp, _i64, _i, _d, _vp, _i, _d, _i, _i, _u64, _u64, _u64, _i, _i,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'svb_villain_cohomology': (_i, [_vp, _vp, _i64, _i, _d, _vp, _i, _u64, _u64, _u64, _i, _vp, _vp, _vp, _vp, _vp]),
    'svb_villain_observables': (_i, [_vp, _i, _vp, _i64, _i, _d, _vp, _vp, _vp]),
    'svb_villain_sweep_tiled': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _d, _vp, _i, _d, _i, _i, _u64, _u64, _u64, _i, _vp, _vp, _vp, _vp]),
    'svb_villain_sweep_tiled_swap': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _d, _vp, _i, _d, _i, _i, _u64, _u64, _u64, _i, _vp, _vp, _vp]),
    'svb_villain_sweep_inplace': (_i, [_vp, _vp, _i64, _i, _d, _vp, _i, _d, _i, _i, _u64, _u64, _u64, _vp, _vp, _vp]),
    'svb_villain_sweep_wavefront': (_i, [_vp, _vp, _i64, _i, _d, _vp, _i, _d, _i, _i, _u64, _u64, _u64, _vp, _vp, _vp, _i64, _vp]),
    'svb_villain_wavefront_workspace': (ctypes.c_longlong, [_i64, _i, _i, _i]),
    'svb_villain_sweep_host': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64, _i, _d, _i, _d, _i, _i, _u64, _u64, _u64, _i, _i, _vp, _i]),
    'svb_worldline_sweep': (_i, [_vp, _vp, _i64, _i, _d, _vp, _i, _i, _i, _i, _u64, _u64, _u64, _i, _i,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'svb_worldline_sweep_overlapped': (_i, [_vp, _vp, _i64, _i, _d, _vp, _i, _i, _i, _u64, _u64, _u64, _vp, _vp,
                                            ctypes.c_uint32, ctypes.c_uint32, _i, _vp]),
    'svb_worldline_observables': (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp]),
    'svb_worldline_wrapping': (_i, [_vp, _vp, _i64, _i, _d, _vp, _i, _i, _u64, _u64, _u64, _i, _vp, _vp, _vp, _vp, _vp]),
    'svb_form_op': (_i, [_i, _i, _i, _vp, _vp, _i64, _i, _vp]),
    'svb_villain_spin_spin': (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    'svb_correlation': (_i, [_i, _vp, _i, _i64, _i, _i, _vp, _vp]),
    'svb_autocorrelation': (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp]),
    'svb_taxicab_correlator': (_i, [_i, _vp, _i64, _i, _d, _vp, _vp, _vp]),
    'svb_block_mean': (_i, [_vp, _vp, _i64, _i64, _i, _i64, _vp, _vp]),
    'svb_bootstrap_mean': (_i, [_vp, _vp, _i64, _i64, _vp, _i, _vp, _vp]),
    'svb_debug_decide_lazy': (_i, [_vp, _vp, _vp, _vp, _i64, ctypes.c_uint32, _u64, _u64, _u64, _vp, _vp, _vp]),
    'svb_philox4x32_10_host': (None, [_vp, _vp, _vp]),
    'svb_villain_draws': (_i, [_i64, _i, _i, _d, _i, _u64, _u64, _u64, _vp, _vp, _vp, _vp]),
}

_lib = None


class SvbError(RuntimeError):
    """A libsvb200 call failed; `.code` is the C return value."""

    def __init__(self, code, message):
        super().__init__(f'libsvb200 error {code}: {message}')
        self.code = code


def load():
    """Load libsvb200.so (once) and attach argument types.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build it with `python -m supervillain_b200.build` '
            '(supervillain_b200 has no CPU fallback).')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code):
    """Map a C return code onto the Python exception types the reference raises."""
    if code == 0:
        return
    msg = load().svb_last_error().decode('utf-8', 'replace')
    if code in (E_PARAM, E_SHAPE, E_NULL, E_ALIGN):
        raise ValueError(msg)
    if code == E_DTYPE:
        raise TypeError(msg)
    if code == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise SvbError(code, msg)
