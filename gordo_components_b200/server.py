"""
The serving side of the hot path without the web framework: what ``POST /gordo/v0/<project>/<name>/anomaly/prediction``
and ``.../prediction`` do between the HTTP layer and ``model.anomaly`` (gordo/server/blueprints/anomaly.py:28-122,
base.py:30-120, utils.py:47-330), as plain functions a Flask / ASGI view can call:

* the wire formats: frames as nested JSON dicts or parquet bytes (``dataframe_to_dict`` / ``dataframe_from_dict`` /
  ``dataframe_into_parquet_bytes`` / ``dataframe_from_parquet_bytes``), and the check of request frames against the model's
  tag list (``verify_dataframe``);
* ``ModelStore``: the models of a project directory kept loaded -- the reference unpickles through ``lru_cache(2)``
  (utils.py:334-353) because a TensorFlow model per machine is heavy; here a model is a few hundred KB of numpy weights whose
  device copy is cached on the estimator, so a whole project stays resident (``max_models`` bounds it if needed);
* ``anomaly_prediction`` / ``prediction``: request payload in, ``Reply(status, body)`` out, with the reference's status
  codes and messages (400 without ``X`` / ``y`` or on unexpected features, 422 when the model is not an anomaly detector).

Many small concurrent requests are better served through ``serving.AnomalyCoalescer`` (one launch for everything that is
waiting); this module is the per-request path and the data formats either way.
"""
import io
import os
import threading
import timeit
from collections import OrderedDict
from typing import Any, Dict, List, Optional, Union

import dateutil.parser
import numpy as np
import pandas as pd

from . import serializer
from .machine.model import utils as model_utils

DELETED_FROM_RESPONSE_COLUMNS = (
    "smooth-tag-anomaly-scaled",
    "smooth-total-anomaly-scaled",
    "smooth-tag-anomaly-unscaled",
    "smooth-total-anomaly-unscaled",
)


# ------------------------------------------------------------------------------------------------ wire formats
def dataframe_into_parquet_bytes(df: pd.DataFrame, compression: str = "snappy") -> bytes:
    import pyarrow as pa
    import pyarrow.parquet as pq

    sink = pa.BufferOutputStream()
    pq.write_table(pa.Table.from_pandas(df), sink, compression=compression)
    return sink.getvalue().to_pybytes()


def dataframe_from_parquet_bytes(buf: bytes) -> pd.DataFrame:
    import pyarrow.parquet as pq

    return pq.read_table(io.BytesIO(buf)).to_pandas()


def dataframe_to_dict(df: pd.DataFrame) -> dict:
    """
    JSON-able form of a frame: ``{column: {index: value}}``, and for two-level columns (the anomaly frame)
    ``{top: {sub: {index: value}}}``; a DatetimeIndex is written as strings (utils.py:86-143).  Built column by column from
    plain lists: going through ``DataFrame.__getitem__`` / ``to_dict`` per top-level name, as the reference does, costs
    tens of milliseconds per response -- more than everything else in a small request together.
    """
    keys = (df.index.astype(str) if isinstance(df.index, pd.DatetimeIndex) else df.index).tolist()
    if not isinstance(df.columns, pd.MultiIndex):
        if not df.columns.is_unique:
            return df.set_axis(keys, axis=0).to_dict()
        return {col: dict(zip(keys, series.tolist())) for col, series in df.items()}
    out: dict = {}
    for (top, sub), series in df.items():
        # a lone column with an empty second level comes out under its own top-level name, as ``DataFrame(series).to_dict()`` gives it
        out.setdefault(top, {})[sub if sub != "" else top] = dict(zip(keys, series.tolist()))
    return out


def blocks_to_dict(index, blocks, columns, skip=()) -> dict:
    """``dataframe_to_dict(frame_from_blocks(index, blocks, columns))`` without building the frame (top-level names in ``skip`` left out)."""
    keys = (index.astype(str) if isinstance(index, pd.DatetimeIndex) else index).tolist()
    out: dict = {}
    names = iter(columns)
    for block in blocks:
        values = block.to_numpy() if isinstance(block, pd.DataFrame) else np.asarray(block)
        for j in range(values.shape[1]):
            top, sub = next(names)
            if top not in skip:
                out.setdefault(top, {})[sub if sub != "" else top] = dict(zip(keys, values[:, j].tolist()))
    return out


def _fast_frame(data: dict) -> Optional[pd.DataFrame]:
    """The frame of a well-formed payload -- every column over the same keys in the same order -- or None (the general path decides)."""
    first = next(iter(data.values()))
    if not isinstance(first, dict) or not first:
        return None
    nested = isinstance(next(iter(first.values())), dict)
    columns = {}
    for top, block in data.items():
        if not isinstance(block, dict):
            return None
        if nested:
            for sub, col in block.items():
                if not isinstance(col, dict):
                    return None
                columns[(top, sub)] = col
        else:
            columns[top] = block
    cols = iter(columns.values())
    keys = list(next(cols))
    if not keys or any(isinstance(v, dict) for v in next(iter(columns.values())).values()):
        return None
    for col in cols:
        if list(col) != keys:
            return None
    index = _parse_keys(keys)
    if index is None:
        return None
    frame = pd.DataFrame({name: list(col.values()) for name, col in columns.items()}, index=index)
    return frame if index.is_monotonic_increasing else frame.sort_index()


def _parse_keys(keys: List[str]) -> Optional[pd.Index]:
    """ISO timestamps of one offset (or none) -> DatetimeIndex; anything else is left to the general path."""
    first = keys[0]
    if not (isinstance(first, str) and len(first) >= 10 and first[4] == "-" and first[7] == "-"):
        return None
    try:
        index = pd.to_datetime(keys, format="ISO8601")
    except (ValueError, TypeError):
        return None
    return index.as_unit("us") if isinstance(index, pd.DatetimeIndex) else None


def dataframe_from_dict(data: dict) -> pd.DataFrame:
    """Inverse of ``dataframe_to_dict``; the index is parsed as ISO timestamps, else as integers, and sorted (utils.py:146-191)."""
    if isinstance(data, dict) and data:
        fast = _fast_frame(data)
        if fast is not None:
            return fast
    if isinstance(data, dict) and any(isinstance(v, dict) for v in data.values()):
        try:
            keys = list(data.keys())
            df = pd.concat((pd.DataFrame.from_dict(data[k]) for k in keys), axis=1, keys=keys)
        except (ValueError, AttributeError):
            df = pd.DataFrame.from_dict(data)
    else:
        df = pd.DataFrame.from_dict(data)
    try:
        df.index = df.index.map(dateutil.parser.isoparse)
    except (TypeError, ValueError):
        df.index = df.index.map(int)
    return df.sort_index()


class Reply:
    """What a view returns: an HTTP status and a JSON-able dict or raw (parquet) bytes."""

    def __init__(self, status: int, body: Union[dict, bytes]):
        self.status, self.body = status, body

    @property
    def content_type(self) -> str:
        return "application/octet-stream" if isinstance(self.body, (bytes, bytearray)) else "application/json"

    def __repr__(self):
        return f"Reply({self.status}, {self.content_type})"


def verify_dataframe(df: pd.DataFrame, expected_columns: List[str]) -> Union[pd.DataFrame, Reply]:
    """
    The request frame reduced / relabelled to the model's tags, or a 400 ``Reply`` (utils.py:206-247): unlabelled frames of
    the right width get the expected names, frames that carry all expected names are reordered, anything else is refused.
    """
    if isinstance(df.columns, pd.MultiIndex):
        return Reply(400, {"message": f"Server does not support multi-level dataframes at this time: {df.columns.tolist()}"})
    if list(df.columns) == list(expected_columns):
        return df
    if all(col in df.columns for col in expected_columns):
        return df[expected_columns]
    if len(df.columns) != len(expected_columns):
        return Reply(400, {"message": f"Unexpected features: was expecting {expected_columns} length of {len(expected_columns)}, "
                                      f"but got {df.columns} length of {len(df.columns)}"})
    df = df.copy(deep=False)
    df.columns = expected_columns
    return df


# ------------------------------------------------------------------------------------------------ resident models
def _tag_names(tags) -> List[str]:
    return [t["name"] if isinstance(t, dict) else str(getattr(t, "name", t)) for t in tags or []]


class ModelStore:
    """
    ``<directory>/<name>/{model.pkl, metadata.json}`` (what ``serializer.dump`` and the builders write) kept loaded.
    Thread safe; ``max_models=None`` keeps everything, otherwise least-recently-used models are dropped.
    """

    def __init__(self, directory: str, max_models: Optional[int] = None):
        self.directory, self.max_models = directory, max_models
        self._models: "OrderedDict[str, Any]" = OrderedDict()
        self._metadata: Dict[str, dict] = {}
        self._lock = threading.Lock()

    def names(self) -> List[str]:
        return sorted(d for d in os.listdir(self.directory) if os.path.isfile(os.path.join(self.directory, d, "model.pkl")))

    def model(self, name: str):
        with self._lock:
            if name in self._models:
                self._models.move_to_end(name)
                return self._models[name]
        path = os.path.join(self.directory, name)
        if not os.path.isfile(os.path.join(path, "model.pkl")):
            raise FileNotFoundError(f"No such model found: '{name}'")
        model = serializer.load(path)
        with self._lock:
            self._models[name] = model
            while self.max_models is not None and len(self._models) > self.max_models:
                self._models.popitem(last=False)
        return model

    def metadata(self, name: str) -> dict:
        with self._lock:
            if name in self._metadata:
                return self._metadata[name]
        meta = serializer.load_metadata(os.path.join(self.directory, name))
        with self._lock:
            self._metadata[name] = meta
        return meta

    def tags(self, name: str) -> List[str]:
        return _tag_names(self.metadata(name).get("dataset", {}).get("tag_list"))

    def target_tags(self, name: str) -> List[str]:
        dataset = self.metadata(name).get("dataset", {})
        return _tag_names(dataset.get("target_tag_list")) or self.tags(name)

    def frequency(self, name: str):
        resolution = self.metadata(name).get("dataset", {}).get("resolution")
        return None if resolution is None else pd.tseries.frequencies.to_offset(resolution)


class ResidentBucket:
    """
    The models of a store that share one feed-forward architecture, served through ONE ``serving.AnomalyCoalescer``: their weights,
    scaler slopes and thresholds sit packed on the device, and whatever requests are waiting -- from any thread, for any of the
    models -- become one fused launch.  Eligible: this package's ``DiffBasedAnomalyDetector`` around a bare ``KerasAutoEncoder``
    (no smoothing window, an affine error scaler); pass ``bucket=`` to ``anomaly_prediction`` and every eligible model is answered
    through it, the rest as before.  The replies are the same bytes either way (rows are independent in the kernel).
    """

    def __init__(self, store: "ModelStore", names: Optional[List[str]] = None, **coalescer_kwargs):
        from . import engine
        from .machine.model.anomaly.diff import _scaler_multiplier
        from .serving import AnomalyCoalescer

        groups: Dict[Any, List[str]] = {}
        for name in names if names is not None else store.names():
            model = store.model(name)
            if self.eligible(model):
                spec = model.base_estimator.model.spec
                has_thr = tuple(t is not None for t in model._thresholds())
                groups.setdefault((tuple(spec.dims), tuple(spec.acts), tuple(spec.l1), has_thr), []).append(name)
        if not groups:
            raise ValueError("no model in the store can be served through a coalescer")
        self.names = max(groups.values(), key=len)  # the largest architecture group
        self.slot = {name: i for i, name in enumerate(self.names)}
        models = [store.model(n) for n in self.names]
        spec = models[0].base_estimator.model.spec
        eng = engine.ff_engine_for(spec)
        torch = engine._torch()
        params = eng.pack_params([m.base_estimator.model.weights for m in models])
        to_dev = lambda rows: torch.from_numpy(np.ascontiguousarray(np.stack(rows), dtype=np.float32)).to(eng.device)  # noqa: E731
        scale = to_dev([_scaler_multiplier(m.scaler, eng.n_out) for m in models])
        feat, agg = zip(*(m._thresholds() for m in models))
        feat_thr = to_dev([np.asarray(f, dtype=np.float32) for f in feat]) if feat[0] is not None else None
        agg_thr = to_dev([np.float32(a) for a in agg]) if agg[0] is not None else None
        self.coalescer = AnomalyCoalescer(eng, params, scale, feat_thr, agg_thr, **coalescer_kwargs)

    @staticmethod
    def eligible(model) -> bool:
        from .machine.model.models import KerasAutoEncoder

        if not (_frame_is_from_blocks(model) and type(model.base_estimator) is KerasAutoEncoder and model.base_estimator.model is not None
                and model.window is None and not (model.require_thresholds and all(t is None for t in model._thresholds()))):
            return False
        from .machine.model.anomaly.diff import _scaler_multiplier

        try:  # a non-affine error scaler (clip=True, QuantileTransformer, ...) is served on the per-request path, not refused for the whole store
            _scaler_multiplier(model.scaler, model.base_estimator.model.spec.dims[-1])
        except (ValueError, AttributeError):
            return False
        return True

    def anomaly_blocks(self, store: "ModelStore", name: str, X: pd.DataFrame, y: pd.DataFrame, frequency=None):
        scores = self.coalescer.anomaly(self.slot[name], X, y)
        return store.model(name).blocks_from_scores(scores, X, y, frequency)

    def close(self):
        self.coalescer.close()


# ------------------------------------------------------------------------------------------------ the two POST views
def _extract_X_y(store: ModelStore, name: str, json: Optional[dict], files: Optional[Dict[str, bytes]]):
    """(X, y) frames of a request -- JSON ``{"X": ..., "y": ...}`` or parquet parts -- or a 400 ``Reply`` (utils.py:250-330)."""
    payload = json if json is not None else (files or {})
    if "X" not in payload:
        return Reply(400, {"message": 'Cannot predict without "X"'})
    load = dataframe_from_dict if json is not None else dataframe_from_parquet_bytes
    X = load(payload["X"])
    y = payload.get("y")
    if y is not None:
        y = load(y)
    tags, targets = store.tags(name), store.target_tags(name)
    X = verify_dataframe(X, tags) if tags else X
    if isinstance(X, Reply):
        return X
    if y is not None and targets:
        y = verify_dataframe(y, targets)
        if isinstance(y, Reply):
            return y
    return X, y


def _frame_is_from_blocks(model) -> bool:
    """True when ``model.anomaly`` is this package's own (not overridden in a subclass): then ``anomaly_blocks`` is the same result."""
    from .machine.model.anomaly.diff import DiffBasedAnomalyDetector

    return isinstance(model, DiffBasedAnomalyDetector) and type(model).anomaly is DiffBasedAnomalyDetector.anomaly \
        and type(model).anomaly_blocks is DiffBasedAnomalyDetector.anomaly_blocks


def _respond(frame: pd.DataFrame, fmt: Optional[str], start: float) -> Reply:
    if fmt == "parquet":
        return Reply(200, dataframe_into_parquet_bytes(frame))
    return Reply(200, {"data": dataframe_to_dict(frame), "time-seconds": f"{timeit.default_timer() - start:.4f}"})


def anomaly_prediction(store: ModelStore, name: str, json: Optional[dict] = None, files: Optional[Dict[str, bytes]] = None,
                       all_columns: bool = False, fmt: Optional[str] = None, bucket: Optional[ResidentBucket] = None) -> Reply:
    """
    ``POST .../<name>/anomaly/prediction`` (anomaly.py:28-122): the anomaly frame of the request's X against its y.  With ``bucket``
    the models it holds are scored through its request coalescer (one launch for everything that is waiting).
    """
    start = timeit.default_timer()
    try:
        model = store.model(name)
    except FileNotFoundError as e:
        return Reply(404, {"message": str(e)})
    xy = _extract_X_y(store, name, json, files)
    if isinstance(xy, Reply):
        return xy
    X, y = xy
    if y is None:
        return Reply(400, {"message": "Cannot perform anomaly without 'y' to compare against."})
    not_a_detector = Reply(422, {"message": f"Model is not an AnomalyDetector, it is of type: {type(model)}"})
    if not hasattr(type(model), "anomaly"):
        return not_a_detector
    skip = () if all_columns else DELETED_FROM_RESPONSE_COLUMNS
    try:
        coalesced = bucket is not None and name in bucket.slot
        if coalesced or (fmt != "parquet" and _frame_is_from_blocks(model)):
            # this package's detectors: straight from the column blocks, no DataFrame in between for JSON
            blocks = bucket.anomaly_blocks(store, name, X, y, store.frequency(name)) if coalesced else model.anomaly_blocks(X, y, frequency=store.frequency(name))
            if fmt != "parquet":
                return Reply(200, {"data": blocks_to_dict(*blocks, skip=skip), "time-seconds": f"{timeit.default_timer() - start:.4f}"})
            frame = model_utils.frame_from_blocks(*blocks)
        else:
            frame = model.anomaly(X, y, frequency=store.frequency(name))
    except AttributeError:  # as the reference: also what a detector without its required thresholds answers (anomaly.py:46-52)
        return not_a_detector
    dropped = [c for c in frame.columns if c[0] in skip]
    if dropped:
        frame = frame.drop(columns=dropped)
    return _respond(frame, fmt, start)


def prediction(store: ModelStore, name: str, json: Optional[dict] = None, files: Optional[Dict[str, bytes]] = None,
               fmt: Optional[str] = None) -> Reply:
    """``POST .../<name>/prediction`` (base.py:30-120): model input and output side by side, no scoring."""
    start = timeit.default_timer()
    try:
        model = store.model(name)
    except FileNotFoundError as e:
        return Reply(404, {"message": str(e)})
    xy = _extract_X_y(store, name, json, files)
    if isinstance(xy, Reply):
        return xy
    X, _ = xy
    try:
        output = model.predict(X) if hasattr(type(model), "predict") or hasattr(model, "predict") else model.transform(X)
    except ValueError as err:
        return Reply(400, {"error": f"ValueError: {err}"})
    except Exception:  # the reference answers every other failure of the model the same way (base.py:83-91)
        return Reply(400, {"error": "Something unexpected happened; check your input data"})
    frame = model_utils.make_base_dataframe(tags=store.tags(name) or list(X.columns), model_input=X.values, model_output=output,
                                            target_tag_list=store.target_tags(name) or None, index=X.index)
    return _respond(frame, fmt, start)
