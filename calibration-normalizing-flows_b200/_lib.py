"""ctypes binding of libcnf_b200.so (the C ABI declared in include/cnf.h).

The library is the product: there is no Python or CPU fallback.  If it has not been built
the import fails loudly with build instructions.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CNF_B200_LIB') or os.path.join(_HERE, 'libcnf_b200.so')     # (override: experiments only)

CNF_MAX_HIDDEN = 4
PREC_FP32, PREC_BF16_TC = 0, 1
HEAD_NLL, HEAD_EXTERNAL = 0, 1
METRICS_PROBS, METRICS_LOGITS, METRICS_CALIBRATED = 0, 1, 2


class FlowDesc(ctypes.Structure):
    _fields_ = [('K', ctypes.c_int32), ('L', ctypes.c_int32), ('n_hidden', ctypes.c_int32),
                ('hidden', ctypes.c_int32 * CNF_MAX_HIDDEN), ('scale', ctypes.c_int32),
                ('shift', ctypes.c_int32), ('precision', ctypes.c_int32),
                ('perm', ctypes.POINTER(ctypes.c_int32))]


class PlanInfo(ctypes.Structure):
    _fields_ = [('n_flat', ctypes.c_int64), ('n_packed', ctypes.c_int64), ('n_tables', ctypes.c_int64),
                ('n_grad_rows', ctypes.c_int64), ('tc_bytes', ctypes.c_int64),
                ('d0', ctypes.c_int32), ('d1', ctypes.c_int32),
                ('hidden_padded', ctypes.c_int32 * CNF_MAX_HIDDEN)]


class CnfError(RuntimeError):
    """A libcnf_b200 call returned a negative CNF_E_* code (``.code``)."""

    def __init__(self, msg, code=0):
        super().__init__(msg)
        self.code = code


_P = ctypes.c_void_p
_DESC = ctypes.POINTER(FlowDesc)
_I64, _I32, _F32 = ctypes.c_int64, ctypes.c_int32, ctypes.c_float

# name -> argtypes; every function returns int except the two noted below.
SIGNATURES = {
    'cnf_plan_info_get': [_DESC, ctypes.POINTER(PlanInfo)],
    'cnf_plan_build': [_DESC, _P, _P],
    'cnf_tc_gather_len': [_DESC, ctypes.POINTER(ctypes.c_int64)],
    'cnf_plan_build_tc': [_DESC, _P],
    'cnf_pack_weights': [_DESC, _P, _P, _P, _P],
    'cnf_pack_weights_tc': [_DESC, _P, _P, _P, _P],
    'cnf_flow_forward': [_DESC, _P, _P, _P, _P, _P, _P, _I64, _P],
    'cnf_flow_inverse': [_DESC, _P, _P, _P, _P, _P, _P, _I64, _P],
    'cnf_flow_predict': [_DESC, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _I32, _P, _P, _P],
    'cnf_flow_apply_host': [_DESC, _P, _P, _P, _P, _P, _I64, _I32, _P, _I64, _I64, _P],
    'cnf_nll_train_step': [_DESC, _P, _P, _P, _P, _I64, _F32, _F32, _F32, _P, _P, _P],
    'cnf_flow_backward': [_DESC, _P, _P, _P, _P, _P, _P, _P, _I64, _P],
    'cnf_grad_reduce': [_DESC, _P, _P, _P, _P],
    'cnf_nll_train_step_rows': [_DESC, _P, _P, _P, _P, _I64, _F32, _F32, _F32, _P, _P, ctypes.POINTER(ctypes.c_int64), _P],
    'cnf_flow_backward_rows': [_DESC, _P, _P, _P, _P, _P, _P, _P, _I64, ctypes.POINTER(ctypes.c_int64), _P],
    'cnf_grad_reduce_rows': [_DESC, _P, _I64, _P, _P, _P],
    'cnf_tc_train_info': [_DESC, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                          ctypes.POINTER(ctypes.c_int64)],
    'cnf_plan_build_tcgrad': [_DESC, _P],
    'cnf_nll_train_step_tc': [_DESC, _P, _P, _P, _P, _I64, _F32, _F32, _F32, _P, _P, _P, _I64,
                              ctypes.POINTER(ctypes.c_int64), _P],
    'cnf_grad_reduce_tc': [_DESC, _P, _I64, _P, _P, _P],
    'cnf_reduce_adam_pack_rows': [_DESC, _P, _I64, _P, _P, _P, _P, _P, _P, _I64, _F32, _F32, _F32, _F32, _P],
    'cnf_fit_full_batch': [_DESC, _P, _P, _P, _P, _I64, _F32, _F32, _F32, _P, _P, _P, _P, _P, _P, _I64, _F32, _F32, _F32, _F32,
                           _I64, _P, _P, _P],
    'cnf_reduce_adam_pack_tc': [_DESC, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _I64, _F32, _F32, _F32, _F32, _P],
    'cnf_adam_step': [_P, _P, _P, _P, _I64, _I64, _F32, _F32, _F32, _F32, _F32, _P],
    'cnf_adam_step_dev': [_P, _P, _P, _P, _I64, _P, _P, _F32, _F32, _F32, _F32, _F32, _P],
    'cnf_sgd_step': [_P, _P, _I64, _F32, _F32, _P],
    'cnf_metrics': [_P, _I32, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P],
    'cnf_calibrated_probs': [_P, _I64, _I32, _P, _P, _P],
    'cnf_affine_const': [_P, _P, _P, _P, _I64, _I32, _I32, _P],
    'cnf_affine_const_backward': [_P, _P, _P, _P, _P, _P, _I64, _I32, _P],
    'cnf_planar_forward': [_P, _P, _P, _P, _P, _P, _I64, _I32, _P],
    'cnf_planar_backward': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P],
    'cnf_radial_forward': [_P, _P, _P, _P, _P, _I64, _I32, _P],
    'cnf_radial_backward': [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P],
}

_lib = None


def load():
    """Load the shared library once; raise if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'cnf_b200: %s not found. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C calibration-normalizing-flows_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.cnf_last_error.restype = ctypes.c_char_p
    lib.cnf_last_error.argtypes = []
    lib.cnf_version.restype = ctypes.c_int
    lib.cnf_version.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CnfError('libcnf_b200 error %d: %s' % (rc, load().cnf_last_error().decode()), rc)


def call(name, *args):
    check(getattr(load(), name)(*args))


def make_desc(K, L, hidden, scale, shift, precision=PREC_FP32, perm=None):
    """Returns (desc, keepalive) -- keepalive holds the perm array the desc points at."""
    d = FlowDesc()
    d.K, d.L, d.n_hidden = int(K), int(L), len(hidden)
    if len(hidden) > CNF_MAX_HIDDEN:
        raise NotImplementedError('cnf_b200 supports at most %d hidden layers per conditioner' % CNF_MAX_HIDDEN)
    for i, h in enumerate(hidden):
        d.hidden[i] = int(h)
    d.scale, d.shift, d.precision = int(bool(scale)), int(bool(shift)), int(precision)
    keep = None
    if perm is not None:
        keep = (ctypes.c_int32 * (L * K))(*[int(v) for row in perm for v in row])
        d.perm = ctypes.cast(keep, ctypes.POINTER(ctypes.c_int32))
    return d, keep
