"""
ctypes binding of the C-ABI library (include/gordo_b200.h, built by csrc/build.py).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is
visible, every compute entry point raises.  PyTorch is used only as the owner of device
memory and streams; the pointers handed to the library are raw device addresses.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgordo_b200.so")

GB_MAX_LAYERS = 16
GB_MAX_WIDTH = 256
ACT_CODES = {"linear": 0, None: 0, "tanh": 1, "relu": 2, "sigmoid": 3}

EXPORTS = (
    "gb_abi_version", "gb_last_error", "gb_device_check", "gb_ffnet_param_count", "gb_ffnet_param_stride",
    "gb_ffae_infer_score", "gb_ffae_tc_supported", "gb_anomaly_score", "gb_anomaly_score_f64", "gb_minmax_fit", "gb_minmax_f64", "gb_thresholds", "gb_thresholds_f64", "gb_cv_moments", "gb_smooth", "gb_quantile", "gb_affine_f64", "gb_ffae_fit_state_stride", "gb_ffae_fit",
    "gb_lstm_param_count", "gb_lstm_param_stride", "gb_lstm_workspace_bytes", "gb_lstm_infer", "gb_lstm_tc_supported", "gb_lstm_tc_workspace_bytes", "gb_lstm_infer_tc", "gb_lstm_fit_workspace_bytes", "gb_lstm_fit",
)


class GbFFNet(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (GB_MAX_LAYERS + 1)), ("act", C.c_int32 * GB_MAX_LAYERS),
                ("l1", C.c_float * GB_MAX_LAYERS)]


class GbJob(C.Structure):
    _fields_ = [("slot", C.c_int32), ("n_rows", C.c_int32), ("x_row", C.c_int64), ("out_row", C.c_int64)]


JOB_DTYPE = np.dtype([("slot", "<i4"), ("n_rows", "<i4"), ("x_row", "<i8"), ("out_row", "<i8")])
assert JOB_DTYPE.itemsize == C.sizeof(GbJob) == 24


class GbFitHParams(C.Structure):
    _fields_ = [("epochs", C.c_int32), ("batch_size", C.c_int32), ("shuffle", C.c_int32), ("l1_div_batch", C.c_int32),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("seed", C.c_uint64), ("step0", C.c_int32), ("reserved", C.c_int32)]


class GbLstmNet(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("n_features", C.c_int32), ("n_features_out", C.c_int32),
                ("units", C.c_int32 * GB_MAX_LAYERS), ("act", C.c_int32 * GB_MAX_LAYERS), ("out_act", C.c_int32),
                ("lookback", C.c_int32)]


class GbLstmFitHParams(C.Structure):
    _fields_ = [("epochs", C.c_int32), ("batch_size", C.c_int32), ("lookahead", C.c_int32), ("primer", C.c_int32),
                ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


class GordoB200Error(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()
_P = C.c_void_p


def _declare(lib):
    lib.gb_abi_version.restype = C.c_int
    lib.gb_last_error.restype = C.c_char_p
    lib.gb_device_check.argtypes = [C.c_int, C.POINTER(C.c_int)]
    for name in ("gb_ffnet_param_count", "gb_ffnet_param_stride", "gb_ffae_fit_state_stride"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.POINTER(GbFFNet)]
    for name in ("gb_lstm_param_count", "gb_lstm_param_stride"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.POINTER(GbLstmNet)]
    lib.gb_lstm_workspace_bytes.restype = C.c_size_t
    lib.gb_lstm_workspace_bytes.argtypes = [C.POINTER(GbLstmNet), C.c_int32, C.c_int32]
    lib.gb_ffae_infer_score.argtypes = [C.POINTER(GbFFNet), _P, _P, C.c_int32, C.c_int32, C.c_int64, C.c_int64] + [_P] * 12 + [C.c_int32, _P]
    lib.gb_ffae_tc_supported.argtypes = [C.POINTER(GbFFNet)]
    lib.gb_ffae_tc_supported.restype = C.c_int
    lib.gb_anomaly_score.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, C.c_int32] + [_P] * 9 + [_P]
    lib.gb_anomaly_score.restype = C.c_int
    lib.gb_anomaly_score_f64.argtypes = lib.gb_anomaly_score.argtypes
    lib.gb_anomaly_score_f64.restype = C.c_int
    lib.gb_minmax_fit.argtypes = [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, _P, C.c_int32, _P]
    lib.gb_minmax_f64.argtypes = [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P]
    lib.gb_minmax_f64.restype = C.c_int
    lib.gb_thresholds.argtypes = [_P, C.c_int32, C.c_int32, _P, _P, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P]
    lib.gb_thresholds_f64.argtypes = lib.gb_thresholds.argtypes
    lib.gb_thresholds_f64.restype = C.c_int
    lib.gb_cv_moments.argtypes = [_P, C.c_int32, _P, _P, C.c_int32, _P, _P]
    lib.gb_cv_moments.restype = C.c_int
    lib.gb_smooth.argtypes = [_P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]
    lib.gb_smooth.restype = C.c_int
    lib.gb_quantile.argtypes = [_P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_float, _P, _P]
    lib.gb_quantile.restype = C.c_int
    lib.gb_affine_f64.argtypes = [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, _P, _P]
    lib.gb_affine_f64.restype = C.c_int
    lib.gb_ffae_fit.argtypes = [C.POINTER(GbFFNet), _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P,
                                C.POINTER(GbFitHParams), _P, _P, _P]
    lib.gb_lstm_infer.argtypes = [C.POINTER(GbLstmNet), _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]
    lib.gb_lstm_tc_supported.argtypes = [C.POINTER(GbLstmNet)]
    lib.gb_lstm_tc_supported.restype = C.c_int
    lib.gb_lstm_tc_workspace_bytes.argtypes = [C.POINTER(GbLstmNet), C.c_int32, C.c_int32, C.c_int32, C.c_int64]
    lib.gb_lstm_tc_workspace_bytes.restype = C.c_size_t
    lib.gb_lstm_infer_tc.argtypes = [C.POINTER(GbLstmNet), _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P, _P, _P]
    lib.gb_lstm_infer_tc.restype = C.c_int
    lib.gb_lstm_fit_workspace_bytes.restype = C.c_size_t
    lib.gb_lstm_fit_workspace_bytes.argtypes = [C.POINTER(GbLstmNet), C.c_int32]
    lib.gb_lstm_fit.argtypes = [C.POINTER(GbLstmNet), _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, C.POINTER(GbLstmFitHParams), _P, _P, _P, _P]
    lib.gb_lstm_fit.restype = C.c_int
    for name in ("gb_device_check", "gb_ffae_infer_score", "gb_ffae_tc_supported", "gb_anomaly_score", "gb_minmax_fit", "gb_thresholds", "gb_smooth", "gb_ffae_fit", "gb_lstm_infer"):
        getattr(lib, name).restype = C.c_int


def load_library():
    """dlopen the C-ABI library (no GPU needed for this step).  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise GordoB200Error(
                    f"{LIB_PATH} is missing: build it with `python gordo_components_b200/csrc/build.py` "
                    "(or __graft_entry__.build()).  gordo_components_b200 has no CPU fallback."
                )
            lib = C.CDLL(LIB_PATH)
            _declare(lib)
            if lib.gb_abi_version() != 2:
                raise GordoB200Error(f"ABI version mismatch: library {lib.gb_abi_version()}, binding 2")
            _lib = lib
    return _lib


_STATUS_EXC = {-1: ValueError, -2: ValueError, -3: ValueError, -4: ValueError, -5: GordoB200Error, -6: GordoB200Error}


def check(rc: int):
    if rc != 0:
        msg = load_library().gb_last_error().decode("utf-8", "replace")
        raise _STATUS_EXC.get(rc, GordoB200Error)(f"gordo_b200 [{rc}]: {msg}")


_device_ok = {}


def require_device(device_index: int = 0) -> int:
    """Fail loudly unless `device_index` is an sm_100 GPU.  Returns its SM count."""
    if device_index in _device_ok:
        return _device_ok[device_index]
    lib = load_library()
    sms = C.c_int(0)
    check(lib.gb_device_check(int(device_index), C.byref(sms)))
    _device_ok[device_index] = sms.value
    return sms.value


def make_ffnet(dims, acts, l1=None) -> GbFFNet:
    n_layers = len(dims) - 1
    if not (1 <= n_layers <= GB_MAX_LAYERS):
        raise ValueError(f"a Dense stack of {n_layers} layers is outside [1, {GB_MAX_LAYERS}]")
    if len(acts) != n_layers:
        raise ValueError("one activation per layer is required")
    net = GbFFNet()
    net.n_layers = n_layers
    for i, d in enumerate(dims):
        net.dims[i] = int(d)
    for i, a in enumerate(acts):
        if a not in ACT_CODES:
            raise ValueError(f"activation {a!r} is not supported by the B200 kernels (supported: tanh, relu, sigmoid, linear)")
        net.act[i] = ACT_CODES[a]
        net.l1[i] = float(l1[i]) if l1 is not None else 0.0
    return net


def make_lstmnet(n_features, units, acts, n_features_out, out_act, lookback) -> GbLstmNet:
    if not (1 <= len(units) <= GB_MAX_LAYERS):
        raise ValueError(f"an LSTM stack of {len(units)} layers is outside [1, {GB_MAX_LAYERS}]")
    net = GbLstmNet()
    net.n_layers = len(units)
    net.n_features, net.n_features_out = int(n_features), int(n_features_out)
    for i, (u, a) in enumerate(zip(units, acts)):
        if a not in ACT_CODES:
            raise ValueError(f"activation {a!r} is not supported by the B200 kernels")
        net.units[i], net.act[i] = int(u), ACT_CODES[a]
    if out_act not in ACT_CODES:
        raise ValueError(f"activation {out_act!r} is not supported by the B200 kernels")
    net.out_act = ACT_CODES[out_act]
    net.lookback = int(lookback)
    return net


def ptr(t):
    """Raw device address of a torch tensor (None -> NULL).  The kernels index dense row-major memory: anything else is refused."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError(f"tensor of shape {tuple(t.shape)} with strides {tuple(t.stride())} is not C-contiguous; the gordo_b200 "
                         "kernels take dense row-major arrays (call .contiguous() / np.ascontiguousarray first)")
    return C.c_void_p(t.data_ptr())
