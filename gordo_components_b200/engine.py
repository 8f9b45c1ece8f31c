"""
Fleet engine: the thin Python layer between gordo-style estimators and the C ABI.

It owns no arithmetic.  PyTorch supplies device memory and the current stream; every
number is produced by a kernel in ``csrc/`` reached through ``_cabi``.  The unit of work is
a *job* (``slot`` = which trained network, a row range of ``x``/``y`` and where the results
go), so one launch scores or trains thousands of machines -- the per-estimator methods in
``machine/model`` are the one-job special case of these functions.

Reference call sites replaced: keras ``Model.predict`` / ``Model.fit`` (gordo/machine/model/
models.py:284,300), sklearn ``MinMaxScaler.fit`` and the pandas arithmetic of
``DiffBasedAnomalyDetector`` (gordo/machine/model/anomaly/diff.py:166-458).
"""
from __future__ import annotations

import ctypes as C
import threading
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _cabi

SCORE_KEYS = ("tag-anomaly-scaled", "tag-anomaly-unscaled", "total-anomaly-scaled", "total-anomaly-unscaled",
              "anomaly-confidence", "total-anomaly-confidence")


def _torch():
    import torch

    return torch


def cuda_device(device=None):
    """Resolve a CUDA device or fail loudly -- there is no CPU path."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise _cabi.GordoB200Error("no CUDA device visible: gordo_components_b200 runs on B200 (sm_100a) only, there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    _cabi.require_device(dev.index)
    return dev


def _stream_ptr():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def make_jobs(slots, n_rows, x_rows, out_rows=None) -> np.ndarray:
    """Structured array of gb_job records."""
    slots = np.asarray(slots)
    jobs = np.zeros(len(slots), dtype=_cabi.JOB_DTYPE)
    jobs["slot"] = slots
    jobs["n_rows"] = n_rows
    jobs["x_row"] = x_rows
    jobs["out_row"] = x_rows if out_rows is None else out_rows
    return jobs


def uniform_jobs(n_machines: int, rows: int) -> np.ndarray:
    """Machine m owns rows [m*rows, (m+1)*rows) and slot m."""
    start = np.arange(n_machines, dtype=np.int64) * rows
    return make_jobs(np.arange(n_machines), rows, start)


def jobs_to_device(jobs: np.ndarray, device):
    torch = _torch()
    raw = torch.from_numpy(np.ascontiguousarray(jobs).view(np.uint8).copy())
    return raw.to(device, non_blocking=False)


class FFEngine:
    """All machines of one Dense-stack architecture (one bucket of the fleet)."""

    def __init__(self, dims: Sequence[int], acts: Sequence[str], l1: Optional[Sequence[float]] = None, device=None):
        self.lib = _cabi.load_library()
        self.device = cuda_device(device)
        self.dims, self.acts = [int(d) for d in dims], list(acts)
        self.l1 = [float(v) for v in (l1 if l1 is not None else [0.0] * (len(dims) - 1))]
        self.net = _cabi.make_ffnet(self.dims, self.acts, self.l1)
        self.n_params = int(self.lib.gb_ffnet_param_count(C.byref(self.net)))
        if self.n_params == 0:
            _cabi.check(-2)
        self.param_stride = int(self.lib.gb_ffnet_param_stride(C.byref(self.net)))
        self.state_stride = int(self.lib.gb_ffae_fit_state_stride(C.byref(self.net)))
        self.n_in, self.n_out = self.dims[0], self.dims[-1]

    # ------------------------------------------------------------------ parameter packing (host side, layout only)
    def pack_params(self, weights_per_slot: Sequence[Sequence[Tuple[np.ndarray, np.ndarray]]]):
        """[(W [in,out], b [out]) per layer] per slot  ->  device tensor [n_slots, param_stride] in Keras order."""
        torch = _torch()
        host = np.zeros((len(weights_per_slot), self.param_stride), dtype=np.float32)
        for s, layers in enumerate(weights_per_slot):
            ofs = 0
            for (W, b), i, o in zip(layers, self.dims[:-1], self.dims[1:]):
                W = np.asarray(W, dtype=np.float32)
                b = np.asarray(b, dtype=np.float32)
                if W.shape != (i, o) or b.shape != (o,):
                    raise ValueError(f"layer weights of shape {W.shape}/{b.shape} do not match the architecture ({i},{o})")
                host[s, ofs : ofs + i * o] = W.ravel()
                ofs += i * o
                host[s, ofs : ofs + o] = b
                ofs += o
        return torch.from_numpy(host).to(self.device)

    def unpack_params(self, params) -> List[List[Tuple[np.ndarray, np.ndarray]]]:
        host = params.detach().cpu().numpy()
        out = []
        for s in range(host.shape[0]):
            ofs, layers = 0, []
            for i, o in zip(self.dims[:-1], self.dims[1:]):
                W = host[s, ofs : ofs + i * o].reshape(i, o).copy()
                ofs += i * o
                b = host[s, ofs : ofs + o].copy()
                ofs += o
                layers.append((W, b))
            out.append(layers)
        return out

    # ------------------------------------------------------------------ K1 + K4
    def infer_score(self, params, jobs_dev, n_jobs: int, max_rows: int, x, y=None, scale=None, feat_thr=None, agg_thr=None,
                    out_rows: Optional[int] = None, want: Sequence[str] = SCORE_KEYS, variant: int = 0, out: Optional[Dict] = None):
        """
        One fused launch: model output (+ the requested anomaly columns) for every job.
        ``want`` selects score outputs (names as in the anomaly frame); outputs are float32 device tensors.
        """
        torch = _torch()
        total = int(out_rows if out_rows is not None else x.shape[0])
        res = out if out is not None else {}

        def buf(name, shape):
            if name not in res:
                res[name] = torch.empty(shape, dtype=torch.float32, device=self.device)
            return res[name]

        o_model = buf("model-output", (total, self.n_out))
        score = y is not None
        sel = set(want) if score else set()
        if scale is None:
            sel -= {"tag-anomaly-scaled", "total-anomaly-scaled", "total-anomaly-confidence"}
        if feat_thr is None:
            sel.discard("anomaly-confidence")
        if agg_thr is None:
            sel.discard("total-anomaly-confidence")
        g = lambda name, shape: buf(name, shape) if name in sel else None  # noqa: E731
        o_ts = g("tag-anomaly-scaled", (total, self.n_out))
        o_tu = g("tag-anomaly-unscaled", (total, self.n_out))
        o_tots = g("total-anomaly-scaled", (total,))
        o_totu = g("total-anomaly-unscaled", (total,))
        o_conf = g("anomaly-confidence", (total, self.n_out))
        o_totc = g("total-anomaly-confidence", (total,))
        p = _cabi.ptr
        _cabi.check(self.lib.gb_ffae_infer_score(
            C.byref(self.net), p(params), p(jobs_dev), int(n_jobs), int(max_rows), int(x.shape[0]), total, p(x), p(y), p(scale), p(feat_thr), p(agg_thr),
            p(o_model), p(o_ts), p(o_tu), p(o_tots), p(o_totu), p(o_conf), p(o_totc), int(variant), _stream_ptr()))
        return res

    # ------------------------------------------------------------------ K2
    def fit(self, params, jobs_dev, n_jobs: int, max_rows: int, x, y, epochs: int = 1, batch_size: int = 32, shuffle=True,
            perm=None, adam: Optional[Dict[str, float]] = None, seed: int = 0, l1_div_batch: bool = False, state=None,
            step0: int = 0):
        """
        Trains every job's slot in place (``params`` is updated).  Returns (loss [n_jobs, epochs], accuracy, (m, v)).
        ``perm`` (int32 [n_jobs, epochs, max_rows]) pins the visiting order (parity tests).
        """
        torch = _torch()
        adam = adam or {}
        hp = _cabi.GbFitHParams()
        hp.epochs, hp.batch_size = int(epochs), int(batch_size)
        hp.shuffle = 2 if perm is not None else (1 if shuffle else 0)
        hp.l1_div_batch = int(bool(l1_div_batch))
        hp.lr, hp.beta1 = float(adam.get("lr", 1e-3)), float(adam.get("beta1", 0.9))
        hp.beta2, hp.eps = float(adam.get("beta2", 0.999)), float(adam.get("eps", 1e-7))
        hp.seed, hp.step0 = int(seed) & (2**64 - 1), int(step0)
        n_slots = params.shape[0]
        if state is None:
            m = torch.zeros((n_slots, self.state_stride), dtype=torch.float32, device=self.device)
            v = torch.zeros_like(m)
        else:
            m, v = state
        loss = torch.empty((n_jobs, epochs), dtype=torch.float32, device=self.device)
        acc = torch.empty((n_jobs, epochs), dtype=torch.float32, device=self.device)
        p = _cabi.ptr
        _cabi.check(self.lib.gb_ffae_fit(C.byref(self.net), p(params), p(m), p(v), p(jobs_dev), int(n_jobs), int(max_rows), p(x), p(y),
                                         p(perm), C.byref(hp), p(loss), p(acc), _stream_ptr()))
        return loss, acc, (m, v)

    # ------------------------------------------------------------------ K7 / K5 / K4-alone (architecture independent)
    def minmax_fit(self, jobs_dev, n_jobs, max_rows, y, n_slots):
        return minmax_fit(jobs_dev, n_jobs, max_rows, y, self.n_out, n_slots, self.device)

    def thresholds(self, jobs_dev, n_jobs, max_rows, tag_unscaled, total_scaled, n_slots, window=6):
        return thresholds(jobs_dev, n_jobs, max_rows, tag_unscaled, total_scaled, self.n_out, n_slots, window, self.device)


def minmax_fit(jobs_dev, n_jobs, max_rows, y, n_out, n_slots, device, return_minmax=False):
    """
    MinMaxScaler.fit per job on device: returns (scale_ [n_slots, n_out], min_ [n_slots, n_out]); with ``return_minmax`` also
    the raw (data_min_, data_max_) the kernel found -- exact float32 values, for callers that need sklearn's float64 arithmetic
    on them (a scaler in front of the network, where ``x * scale_ + min_`` cancels for offset-dominated tags).
    """
    torch = _torch()
    lib = _cabi.load_library()
    scale = torch.ones((n_slots, n_out), dtype=torch.float32, device=device)
    offset = torch.zeros((n_slots, n_out), dtype=torch.float32, device=device)
    ws = torch.empty((n_slots, 2, n_out), dtype=torch.float32, device=device)
    p = _cabi.ptr
    _cabi.check(lib.gb_minmax_fit(p(jobs_dev), int(n_jobs), int(max_rows), p(y), int(n_out), p(scale), p(offset), p(ws), int(n_slots), _stream_ptr()))
    if return_minmax:
        return scale, offset, ws[:, 0], ws[:, 1]
    return scale, offset


def minmax_f64(jobs_dev, n_jobs, max_rows, y64, n_slots):
    """Column (min, max) of float64 targets per job: two [n_slots, n_out] float64 tensors (NaNs skipped, +-inf when a column has none)."""
    torch = _torch()
    lib = _cabi.load_library()
    if y64.dtype != torch.float64:
        raise ValueError(f"minmax_f64 takes float64 targets, got {y64.dtype}")
    n_out = y64.shape[1]
    mm = torch.empty((int(n_slots), 2, n_out), dtype=torch.float64, device=y64.device)
    p = _cabi.ptr
    _cabi.check(lib.gb_minmax_f64(p(jobs_dev), int(n_jobs), int(max_rows), p(y64), int(n_out), p(mm), int(n_slots), _stream_ptr()))
    return mm[:, 0], mm[:, 1]


def thresholds(jobs_dev, n_jobs, max_rows, tag_unscaled, total_scaled, n_out, n_slots, window, device):
    """
    rolling(window).min().max() per tag and for the aggregate series: (feat_thr [n_slots, n_out], agg_thr [n_slots]).
    float32 or float64 score arrays (the dtype of ``tag_unscaled`` decides; both arrays must agree).
    """
    torch = _torch()
    lib = _cabi.load_library()
    dtype = tag_unscaled.dtype
    if total_scaled.dtype != dtype or dtype not in (torch.float32, torch.float64):
        raise ValueError(f"thresholds need two float32 or two float64 arrays, got {dtype} / {total_scaled.dtype}")
    feat = torch.full((n_slots, n_out), float("nan"), dtype=dtype, device=device)
    agg = torch.full((n_slots,), float("nan"), dtype=dtype, device=device)
    p = _cabi.ptr
    fn = lib.gb_thresholds if dtype == torch.float32 else lib.gb_thresholds_f64
    _cabi.check(fn(p(jobs_dev), int(n_jobs), int(max_rows), p(tag_unscaled), p(total_scaled), int(n_out), int(window),
                                  p(feat), p(agg), int(n_slots), _stream_ptr()))
    return feat, agg


def cv_moments(jobs_dev, n_jobs, yhat, y, n_out):
    """Per job and column: sums of e, e^2, |e|, (y-y0), (y-y0)^2 over the job's rows (e = yhat - y): [n_jobs, 5, n_out] float64."""
    torch = _torch()
    lib = _cabi.load_library()
    out = torch.zeros((int(n_jobs), 5, int(n_out)), dtype=torch.float64, device=yhat.device)
    if int(n_jobs) == 0:
        return out
    p = _cabi.ptr
    _cabi.check(lib.gb_cv_moments(p(jobs_dev), int(n_jobs), p(yhat), p(y), int(n_out), p(out), _stream_ptr()))
    return out


def anomaly_score(jobs_dev, n_jobs, max_rows, yhat, y, n_out, scale=None, feat_thr=None, agg_thr=None, want=SCORE_KEYS, device=None):
    """
    Anomaly columns for predictions that already exist (base estimators that are not ours, LSTM outputs).  float64 ``yhat`` selects
    the float64 kernel (the reference's own precision, diff.py:268-300, 350-385): every operand must then be float64 and so are the results.
    """
    torch = _torch()
    lib = _cabi.load_library()
    device = yhat.device
    total = yhat.shape[0]
    dtype = yhat.dtype
    for name, t in (("y", y), ("scale", scale), ("feat_thr", feat_thr), ("agg_thr", agg_thr)):
        if t is not None and t.dtype != dtype:
            raise ValueError(f"anomaly_score: {name} is {t.dtype}, yhat is {dtype}")
    fn = {torch.float32: lib.gb_anomaly_score, torch.float64: lib.gb_anomaly_score_f64}.get(dtype)
    if fn is None:
        raise ValueError(f"anomaly_score takes float32 or float64 arrays, not {dtype}")
    sel = set(want)
    if scale is None:
        sel -= {"tag-anomaly-scaled", "total-anomaly-scaled", "total-anomaly-confidence"}
    if feat_thr is None:
        sel.discard("anomaly-confidence")
    if agg_thr is None:
        sel.discard("total-anomaly-confidence")
    res = {}

    def g(name, shape):
        if name in sel:
            res[name] = torch.empty(shape, dtype=dtype, device=device)
            return res[name]
        return None

    o_ts = g("tag-anomaly-scaled", (total, n_out))
    o_tu = g("tag-anomaly-unscaled", (total, n_out))
    o_tots = g("total-anomaly-scaled", (total,))
    o_totu = g("total-anomaly-unscaled", (total,))
    o_conf = g("anomaly-confidence", (total, n_out))
    o_totc = g("total-anomaly-confidence", (total,))
    p = _cabi.ptr
    _cabi.check(fn(p(jobs_dev), int(n_jobs), int(max_rows), p(yhat), p(y), int(n_out), p(scale), p(feat_thr), p(agg_thr),
                                     p(o_ts), p(o_tu), p(o_tots), p(o_totu), p(o_conf), p(o_totc), _stream_ptr()))
    return res


def quantile(jobs_dev, n_jobs, max_rows, arr, q: float):
    """pandas ``.quantile(q)`` (linear interpolation, NaNs skipped) of every column of ``arr`` per job -> [n_jobs, n_cols]."""
    torch = _torch()
    lib = _cabi.load_library()
    a2 = arr if arr.dim() == 2 else arr.reshape(-1, 1)
    out = torch.empty((int(n_jobs), a2.shape[1]), dtype=torch.float32, device=a2.device)
    p = _cabi.ptr
    _cabi.check(lib.gb_quantile(p(jobs_dev), int(n_jobs), int(max_rows), p(a2), int(a2.shape[1]), float(q), p(out), _stream_ptr()))
    return out


def affine_f64(jobs_dev, n_jobs, max_rows, x64, a, b, out_rows=None):
    """Per-feature ``x * a[slot] + b[slot]`` in float64 on the device, rounded once to float32 (sklearn scalers' transform)."""
    torch = _torch()
    lib = _cabi.load_library()
    n_cols = x64.shape[1]
    out = torch.empty((int(out_rows if out_rows is not None else x64.shape[0]), n_cols), dtype=torch.float32, device=x64.device)
    p = _cabi.ptr
    _cabi.check(lib.gb_affine_f64(p(jobs_dev), int(n_jobs), int(max_rows), p(x64), int(n_cols), p(a), p(b), p(out), _stream_ptr()))
    return out


SMOOTH_METHODS = {"smm": 0, "sma": 1, "ewma": 2}


def smooth(jobs_dev, n_jobs, arr, window: int, method: str, max_rows: Optional[int] = None):
    """
    Rolling median / mean / EWMA of every column of ``arr`` ([rows] or [rows, cols]) per job (pandas semantics, NaNs included).
    ``max_rows``: longest job (defaults to the length of ``arr``, always an upper bound).
    """
    torch = _torch()
    lib = _cabi.load_library()
    if method not in SMOOTH_METHODS:
        raise ValueError(f"smoothing_method {method!r} must be one of {sorted(SMOOTH_METHODS)}")
    n_cols = 1 if arr.dim() == 1 else arr.shape[1]
    out = torch.full_like(arr, float("nan"))
    p = _cabi.ptr
    _cabi.check(lib.gb_smooth(p(jobs_dev), int(n_jobs), int(max_rows if max_rows is not None else arr.shape[0]), p(arr), int(n_cols), int(window),
                              SMOOTH_METHODS[method], p(out), _stream_ptr()))
    return out


class LSTMEngine:
    """All machines of one LSTM-stack architecture."""

    def __init__(self, n_features, units, acts, n_features_out, out_func, lookback, device=None):
        self.lib = _cabi.load_library()
        self.device = cuda_device(device)
        self.n_features, self.units, self.acts = int(n_features), [int(u) for u in units], list(acts)
        self.n_out, self.out_func, self.lookback = int(n_features_out), out_func, int(lookback)
        self.net = _cabi.make_lstmnet(n_features, units, acts, n_features_out, out_func, lookback)
        self.n_params = int(self.lib.gb_lstm_param_count(C.byref(self.net)))
        if self.n_params == 0:
            _cabi.check(-2)
        self.param_stride = int(self.lib.gb_lstm_param_stride(C.byref(self.net)))

    def pack_params(self, weights_per_slot):
        """([(kernel [in,4u], recurrent [u,4u], bias [4u]) per layer], (Wd, bd)) per slot -> device [n_slots, stride]."""
        torch = _torch()
        host = np.zeros((len(weights_per_slot), self.param_stride), dtype=np.float32)
        for s, (layers, (Wd, bd)) in enumerate(weights_per_slot):
            flat = []
            for K, U, b in layers:
                flat += [np.asarray(K, np.float32).ravel(), np.asarray(U, np.float32).ravel(), np.asarray(b, np.float32).ravel()]
            flat += [np.asarray(Wd, np.float32).ravel(), np.asarray(bd, np.float32).ravel()]
            vec = np.concatenate(flat)
            if vec.size != self.n_params:
                raise ValueError(f"LSTM weights hold {vec.size} values, architecture needs {self.n_params}")
            host[s, : vec.size] = vec
        return torch.from_numpy(host).to(self.device)

    def unpack_params(self, params):
        """device [n_slots, stride] -> per slot ([(kernel, recurrent, bias) per layer], (Wd, bd)) as host arrays."""
        host = params.detach().cpu().numpy()
        out = []
        for vec in host:
            layers, ofs, i = [], 0, self.n_features
            for u in self.units:
                K = vec[ofs:ofs + i * 4 * u].reshape(i, 4 * u).copy(); ofs += i * 4 * u
                U = vec[ofs:ofs + u * 4 * u].reshape(u, 4 * u).copy(); ofs += u * 4 * u
                b = vec[ofs:ofs + 4 * u].copy(); ofs += 4 * u
                layers.append((K, U, b))
                i = u
            Wd = vec[ofs:ofs + i * self.n_out].reshape(i, self.n_out).copy(); ofs += i * self.n_out
            out.append((layers, (Wd, vec[ofs:ofs + self.n_out].copy())))
        return out

    def fit(self, params, jobs_dev, n_jobs, max_windows, x, y, epochs: int = 1, batch_size: int = 32, lookahead: int = 0,
            primer: bool = True, adam: Optional[Dict[str, float]] = None, state=None):
        """
        Trains every job's slot in place by back-propagation through time (``jobs`` count windows; target of window j is
        y[x_row + j + lookback - 1 + lookahead]).  Returns (loss [n_jobs, epochs], accuracy, (m, v, t)).
        """
        torch = _torch()
        adam = adam or {}
        hp = _cabi.GbLstmFitHParams()
        hp.epochs, hp.batch_size, hp.lookahead, hp.primer = int(epochs), int(batch_size), int(lookahead), int(bool(primer))
        hp.lr, hp.beta1 = float(adam.get("lr", 1e-3)), float(adam.get("beta1", 0.9))
        hp.beta2, hp.eps = float(adam.get("beta2", 0.999)), float(adam.get("eps", 1e-7))
        if state is None:
            m = torch.zeros_like(params)
            v = torch.zeros_like(params)
            t = torch.zeros((params.shape[0],), dtype=torch.int32, device=self.device)
        else:
            m, v, t = state
        ws_bytes = int(self.lib.gb_lstm_fit_workspace_bytes(C.byref(self.net), int(n_jobs)))
        ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=self.device)
        loss = torch.zeros((n_jobs, max(epochs, 1)), dtype=torch.float32, device=self.device)[:, :epochs]
        acc = torch.zeros((n_jobs, max(epochs, 1)), dtype=torch.float32, device=self.device)[:, :epochs]
        loss, acc = loss.contiguous(), acc.contiguous()
        p = _cabi.ptr
        _cabi.check(self.lib.gb_lstm_fit(C.byref(self.net), p(params), p(m), p(v), p(t), p(jobs_dev), int(n_jobs), int(max_windows), p(x), p(y),
                                         C.byref(hp), p(ws), p(loss), p(acc), _stream_ptr()))
        return loss, acc, (m, v, t)

    @property
    def tc_supported(self) -> bool:
        return self.lib.gb_lstm_tc_supported(C.byref(self.net)) == 0

    def infer(self, params, jobs_dev, n_jobs, max_windows, x, out_rows, variant: int = 0):
        """
        out[j] = net(x[j : j + lookback]) for every job's windows (jobs' n_rows counts windows).
        variant 0 = tcgen05 kernel when the layer widths allow it, 1 = fp32 CUDA-core kernel, 2 = tcgen05 (error if unsupported).
        """
        torch = _torch()
        out = torch.empty((int(out_rows), self.n_out), dtype=torch.float32, device=self.device)
        p = _cabi.ptr
        if variant == 2 or (variant == 0 and self.tc_supported):
            ws_bytes = int(self.lib.gb_lstm_tc_workspace_bytes(C.byref(self.net), int(params.shape[0]), int(n_jobs), int(max_windows), int(x.shape[0])))
            if ws_bytes == 0:
                _cabi.check(self.lib.gb_lstm_tc_supported(C.byref(self.net)))
            ws = torch.empty((ws_bytes + 255,), dtype=torch.uint8, device=self.device)
            _cabi.check(self.lib.gb_lstm_infer_tc(C.byref(self.net), p(params), int(params.shape[0]), p(jobs_dev), int(n_jobs), int(max_windows), p(x),
                                                  int(x.shape[0]), p(out), p(ws), _stream_ptr()))
            return out
        _cabi.check(self.lib.gb_lstm_infer(C.byref(self.net), p(params), p(jobs_dev), int(n_jobs), int(max_windows), p(x), p(out), None, _stream_ptr()))
        return out


_ff_engines: Dict[tuple, FFEngine] = {}
_lstm_engines: Dict[tuple, LSTMEngine] = {}
_engine_lock = threading.Lock()  # gordo.server calls into shared models from several gunicorn threads


def ff_engine_for(spec, device=None) -> FFEngine:
    dev = cuda_device(device)
    key = (spec.key(), tuple(spec.l1), dev.index)
    with _engine_lock:
        if key not in _ff_engines:
            _ff_engines[key] = FFEngine(spec.dims, spec.acts, spec.l1, dev)
        return _ff_engines[key]


def lstm_engine_for(spec, device=None) -> LSTMEngine:
    dev = cuda_device(device)
    key = (spec.key(), dev.index)
    with _engine_lock:
        if key not in _lstm_engines:
            _lstm_engines[key] = LSTMEngine(spec.n_features, spec.lstm_units, spec.acts, spec.n_features_out, spec.out_func, spec.lookback_window, dev)
        return _lstm_engines[key]


def to_device_f32(a, device):
    """Host array/frame -> contiguous float32 device tensor (the reference casts to floatx=float32 too [3P scikeras])."""
    torch = _torch()
    arr = np.ascontiguousarray(np.asarray(getattr(a, "values", a), dtype=np.float32))
    if arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    return torch.from_numpy(arr).to(device)
