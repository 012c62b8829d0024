"""ctypes binding of librbslam.so (include/rbslam.h).

This is the Python twin of the MEX gateway: it marshals NumPy arrays (made
column-major fp64, exactly what ``mxGetDoubles`` would hand over) into the C ABI.
There is no fallback: if the shared library is missing or fails to load, import
of the compute entry points raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "librbslam.so")

OK, EARG, ECUDA, ENOTPD, EMODEL = 0, 1, 2, 3, 4
MODEL_DENSE_MAG3D, MODEL_DENSE_RADIO2D, MODEL_SPARSE_VISUAL2D = 1, 2, 3
RNG_INJECTED, RNG_PHILOX = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32), ("model", C.c_int32),
        ("N", C.c_int32), ("T", C.c_int32), ("m_basis", C.c_int32),
        ("NN", c_int32_p), ("L", c_double_p),
        ("cam_f", C.c_double), ("cam_fp", C.c_double), ("cam_fw", C.c_double),
        ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("ld", C.c_int32),
        ("information_form", C.c_int32), ("keep_history", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32), ("kalman_variant", C.c_int32),
    ]


class Inputs(C.Structure):
    _fields_ = [
        ("T", C.c_int32), ("odometry", c_double_p), ("odo_rows", C.c_int32),
        ("y", c_double_p), ("x0_nonLin", c_double_p), ("x0_lin", c_double_p),
        ("x0_lin_cols", C.c_int32), ("P0_lin", c_double_p), ("Q", c_double_p),
        ("Q_pages", C.c_int32), ("R", c_double_p), ("dt", c_double_p), ("dt_len", C.c_int32),
        ("U", c_double_p), ("Z", c_double_p), ("Uend", c_double_p),
        ("forced_ancestors", c_int32_p), ("forced_ak", c_int32_p),
    ]


class FilterOutputs(C.Structure):
    _fields_ = [
        ("traj_max", c_double_p), ("traj_mean", c_double_p), ("xl_max", c_double_p),
        ("xl_mean", c_double_p), ("P_max", c_double_p), ("P_mean", c_double_p),
        ("traj_sample_iwmax", c_double_p), ("xn_traj", c_double_p),
        ("logw_hist", c_double_p), ("w_hist", c_double_p), ("ancestors", c_int32_p),
    ]


class SmootherOutputs(C.Structure):
    _fields_ = [("XNK", c_double_p), ("XLK", c_double_p), ("PK", c_double_p),
                ("AI", c_double_p), ("ak", c_int32_p)]


STEP_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.c_int32)

# every symbol include/rbslam.h declares: name -> (restype, argtypes)
_ctx = C.c_void_p
SYMBOLS = {
    "rbslam_version": (C.c_int, []),
    "rbslam_device_count": (C.c_int, []),
    "rbslam_create": (C.c_int, [C.POINTER(_ctx), C.POINTER(Config)]),
    "rbslam_create_group": (C.c_int, [C.POINTER(_ctx), C.POINTER(Config), c_int32_p, C.c_int32]),
    "rbslam_create_replicas": (C.c_int, [C.POINTER(_ctx), C.POINTER(Config), c_int32_p, C.c_int32]),
    "rbslam_destroy": (None, [_ctx]),
    "rbslam_last_error": (C.c_char_p, [_ctx]),
    "rbslam_dims": (C.c_int, [_ctx, c_int32_p]),
    "rbslam_filter_run": (C.c_int, [_ctx, C.POINTER(Inputs), C.POINTER(FilterOutputs)]),
    "rbslam_smoother_run": (C.c_int, [_ctx, C.POINTER(Inputs), C.c_int32, C.c_int32,
                                      C.POINTER(SmootherOutputs)]),
    "rbslam_step_callback": (C.c_int, [_ctx, STEP_FN, C.c_void_p]),
    "rbslam_filter_begin": (C.c_int, [_ctx, C.POINTER(Inputs)]),
    "rbslam_filter_step": (C.c_int, [_ctx]),
    "rbslam_filter_end": (C.c_int, [_ctx, C.POINTER(FilterOutputs)]),
    "rbslam_sync": (C.c_int, [_ctx]),
    "rbslam_read_particles": (C.c_int, [_ctx, c_double_p, c_double_p, c_double_p, c_double_p,
                                        c_double_p, c_int32_p]),
    "rbslam_read_trajectories": (C.c_int, [_ctx, c_double_p, c_double_p, c_double_p, c_double_p]),
    "rbslam_read_information": (C.c_int, [_ctx, c_double_p, c_double_p, c_double_p]),
    "rbslam_counters": (C.c_int, [_ctx, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64)]),
    "rbslam_status_counters": (C.c_int, [_ctx, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rbslam_event_record": (C.c_int, [_ctx, C.c_int32]),
    "rbslam_event_elapsed": (C.c_int, [_ctx, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "rbslam_phase_timing": (C.c_int, [_ctx, C.c_int32]),
    "rbslam_phase_times": (C.c_int, [_ctx, c_double_p]),
    "rbslam_stream": (C.c_void_p, [_ctx]),
    "rbslam_op_resample": (C.c_int, [_ctx, C.c_int32, c_double_p, C.c_int32, c_double_p, c_int32_p]),
    "rbslam_op_normalize": (C.c_int, [_ctx, C.c_int32, c_double_p, c_double_p, c_int32_p]),
    "rbslam_op_propagate": (C.c_int, [_ctx, C.c_int32, c_double_p, c_int32_p, c_double_p, C.c_double,
                                      c_double_p, c_double_p, c_double_p]),
    "rbslam_op_meas_jacobian": (C.c_int, [_ctx, C.c_int32, c_double_p, c_double_p, c_double_p,
                                          c_double_p]),
    "rbslam_op_jacobian_phi3d": (C.c_int, [_ctx, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p]),
    "rbslam_op_kalman_update": (C.c_int, [_ctx, C.c_int32, c_double_p, c_double_p, c_double_p,
                                          c_double_p, C.c_double, c_double_p, c_double_p, c_double_p]),
    "rbslam_op_dyn_logweight": (C.c_int, [_ctx, C.c_int32, c_double_p, c_double_p, c_double_p,
                                          C.c_double, c_double_p, C.c_int32, c_double_p]),
    "rbslam_op_ancestor_weights": (C.c_int, [_ctx, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_double_p,
                                             c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                             C.c_double, c_double_p]),
    "rbslam_ekf_run": (C.c_int, [_ctx, C.c_int32, c_double_p, C.c_int32, c_double_p, c_double_p, c_double_p,
                                 c_double_p, c_double_p, C.c_int32, c_double_p, c_double_p, C.c_int32, c_double_p,
                                 c_double_p, c_double_p, c_double_p, c_double_p]),
    "rbslam_localization_run": (C.c_int, [_ctx, C.c_int32, C.c_int32, c_double_p, C.c_int32, c_double_p, c_double_p,
                                          C.c_int32, c_double_p, C.c_int32, c_double_p, C.c_int32, c_double_p,
                                          c_double_p, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p,
                                          c_double_p, c_int32_p, c_double_p, c_int32_p]),
    "rbslam_plan_migration": (C.c_int, [C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int32_p,
                                        c_int32_p]),
    "rbslam_plan_shard": (C.c_int, [C.c_int32, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                    c_int32_p, c_int32_p]),
    "rbslam_op_plan_shard": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [c_int32_p] * 11),
    "rbslam_ipc_count": (C.c_int, []),
    "rbslam_ipc_export": (C.c_int, [_ctx, C.c_int32, C.c_void_p]),
    "rbslam_ipc_import": (C.c_int, [_ctx, C.c_int32, C.c_int32, C.c_void_p]),
}

_lib = None


class RbslamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rbslam error %d: %s" % (code, msg))
        self.code = code


class UnsupportedModelError(RbslamError):
    """rbslam:unsupportedModel -- the handle is not one of the registered families."""


def lib():
    """Load librbslam.so (once).  Raises if it is missing: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "librbslam.so not found at %s; build it with "
                "`python rao-blackwellized-slam-smoothing_b200/build.py` "
                "(there is no CPU fallback by design)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def iptr(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def fcol(a, dtype=np.float64):
    """Column-major contiguous copy/view, the layout MATLAB arrays have."""
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def check(ctx, rc):
    if rc != OK:
        msg = lib().rbslam_last_error(ctx)
        msg = msg.decode() if msg else ""
        if rc == EMODEL:
            raise UnsupportedModelError(rc, msg)
        raise RbslamError(rc, msg)
