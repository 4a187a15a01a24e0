"""ctypes loader for libgokalman_b200.so (the C-ABI in include/gokalman_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C gokalman_b200/csrc`.  There is
no fallback: if it is missing, or no sm_100 device is usable, every call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgokalman_b200.so")

VANILLA, PREDICTOR, INFORMATION, SQRT, HYBRID, SRIF = range(6)
HOST, DEVICE = 0, 1
NOISE_PHILOX, NOISE_REPLAY = 0, 1
REDUCE_NCCL, REDUCE_PEER = 0, 1
F_MEAS, F_EKF, F_SNC = 1, 2, 4

STATUS_NAMES = {
    0: "ok", -1: "dimensions must agree", -2: "could not invert H*P*H' + R", -3: "matrix is not symmetric",
    -4: "kf is locked (call Prepare() first)", -5: "could not invert Phi", -6: "cannot invert R",
    -7: "no noise defined at this step", -8: "unsupported shape", -9: "CUDA error", -10: "bad argument",
    -11: "non-finite estimate",
}


class GkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gokalman_b200 error %d (%s): %s" % (code, STATUS_NAMES.get(code, "?"), msg))
        self.code = code


class Outputs(C.Structure):
    _fields_ = [("mem", C.c_int), ("every_step", C.c_int),
                ("state", C.c_void_p), ("meas", C.c_void_p), ("innov", C.c_void_p), ("covar", C.c_void_p),
                ("pred_covar", C.c_void_p), ("gain", C.c_void_p), ("obs_dev", C.c_void_p), ("status", C.c_void_p)]


class McConfig(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_int), ("m", C.c_int), ("c", C.c_int),
                ("F", C.c_void_p), ("G", C.c_void_p), ("H", C.c_void_p), ("Q", C.c_void_p), ("R", C.c_void_p),
                ("x0_truth", C.c_void_p), ("x0_filter", C.c_void_p), ("P0", C.c_void_p),
                ("trials", C.c_int64), ("trial_offset", C.c_int64), ("steps", C.c_int),
                ("controls", C.c_void_p), ("noise_mode", C.c_int), ("seed", C.c_uint64),
                ("w", C.c_void_p), ("v", C.c_void_p), ("noise_mem", C.c_int),
                ("with_nees", C.c_int), ("with_nis", C.c_int), ("info_raw_init", C.c_int), ("device", C.c_int),
                ("filter_F", C.c_void_p), ("filter_G", C.c_void_p), ("filter_H", C.c_void_p), ("filter_Q", C.c_void_p),
                ("filter_R", C.c_void_p)]


class OdConfig(C.Structure):
    _fields_ = [("mu", C.c_double), ("j2", C.c_double), ("re", C.c_double), ("dt", C.c_double),
                ("orbit0", C.c_void_p), ("orbit_mem", C.c_int), ("station", C.c_void_p), ("truth_obs", C.c_void_p),
                ("sigma_range", C.c_double), ("sigma_rate", C.c_double), ("seed", C.c_uint64), ("filter_offset", C.c_int64)]


class McOutputs(C.Structure):
    _fields_ = [("mem", C.c_int), ("sums_only", C.c_int),
                ("nis", C.c_void_p), ("nees", C.c_void_p), ("sum_d", C.c_void_p), ("sum_dd", C.c_void_p), ("x_ref", C.c_void_p),
                ("truth_x", C.c_void_p), ("truth_y", C.c_void_p), ("noise_w", C.c_void_p), ("noise_v", C.c_void_p),
                ("status", C.c_void_p), ("first_error", C.c_void_p)]


# every symbol include/gokalman_b200.h declares: (name, restype, argtypes)
_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64
SYMBOLS = [
    ("gkb_version", C.c_char_p, []),
    ("gkb_last_error", C.c_char_p, []),
    ("gkb_device_count", _i, []),
    ("gkb_shape_supported", _i, [_i, _i, _i]),
    ("gkb_create_lti", _i, [_i, _i, _i, _i, _i64, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    ("gkb_create_information_from_state", _i, [_i, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    ("gkb_create_hybrid", _i, [_i, _i, _i, _i64, _i, _vp, _i, _vp, _vp, _vp, C.POINTER(_vp)]),
    ("gkb_create_srif", _i, [_i, _i, _i64, _i, _vp, _i, _vp, _vp, _i, C.POINTER(_vp)]),
    ("gkb_destroy", None, [_vp]),
    ("gkb_set_state_transition", _i, [_vp, _vp]),
    ("gkb_set_input_control", _i, [_vp, _i, _vp]),
    ("gkb_set_measurement_matrix", _i, [_vp, _i, _vp]),
    ("gkb_set_noise", _i, [_vp, _vp, _i, _vp]),
    ("gkb_set_replay_noise", _i, [_vp, _i, _vp, _vp, _i]),
    ("gkb_set_philox_noise", _i, [_vp, C.c_uint64, _i64]),
    ("gkb_awgn_sample", _i, [_i, _i, _vp, _vp, C.c_uint64, _i64, _i, _i, _vp, _vp, _vp]),
    ("gkb_reset", _i, [_vp]),
    ("gkb_set_stream", _i, [_vp, _vp]),
    ("gkb_set_strict", _i, [_vp, _i]),
    ("gkb_n_filters", _i64, [_vp]),
    ("gkb_filter_major", _i, [_vp]),
    ("gkb_step", _i, [_vp]),
    ("gkb_update", _i, [_vp, _i, _vp, _i, _vp, _i, C.POINTER(Outputs)]),
    ("gkb_nl_run", _i, [_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i, C.POINTER(Outputs)]),
    ("gkb_od_synthesize", _i, [C.POINTER(OdConfig), _i, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    ("gkb_od_run", _i, [_vp, C.POINTER(OdConfig), _i, _vp, C.POINTER(Outputs)]),
    ("gkb_van_loan", _i, [_i, _i, _i64, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    ("gkb_get_state", _i, [_vp, _vp, _vp]),
    ("gkb_set_state", _i, [_vp, _vp, _vp]),
    ("gkb_smooth_all", _i, [_i, _i, _i64, _i, _vp, _i, _vp, _vp, _i, _vp]),
    ("gkb_householder_transf", _i, [_i, _i, _i64, _i, _vp, _i]),
    ("gkb_batch_solve", _i, [_i, _i, _i, _i64, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    ("gkb_mc_chisquare", _i, [C.POINTER(McConfig), C.POINTER(McOutputs)]),
    ("gkb_mc_chisquare_multi", _i, [C.POINTER(McConfig), C.POINTER(_i), _i, _i, C.POINTER(McOutputs)]),
    ("gkb_last_kernel_ms", C.c_float, []),
    ("gkb_last_main_kernel_ms", C.c_float, []),
    ("gkb_last_kernel_launches", _i, []),
]

_lib = None


def load():
    """Load the CUDA engine.  Raises (never falls back) when the shared library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gokalman_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GkbError(rc, load().gkb_last_error().decode("utf-8", "replace"))
