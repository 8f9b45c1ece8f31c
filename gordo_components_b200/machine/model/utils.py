"""
Frame assembly and metric helpers with the contract of gordo/machine/model/utils.py
(metric_wrapper :18-46, make_base_dataframe :49-165).
"""
from __future__ import annotations

import functools
from datetime import timedelta
from typing import List, Optional, Sequence, Union

import numpy as np
import pandas as pd


def _tag_name(tag) -> str:
    return str(getattr(tag, "name", tag))


def metric_wrapper(metric, scaler=None):
    """
    Make ``metric(y_true, y_pred)`` robust to models whose output is shorter than the target (LSTM offset):
    ``y_true`` is tail-aligned to ``y_pred``; if a fitted ``scaler`` is given both are transformed first.
    """

    @functools.wraps(metric)
    def _wrapper(y_true, y_pred, *args, **kwargs):
        if scaler:
            y_true = scaler.transform(y_true)
            y_pred = scaler.transform(y_pred)
        return metric(y_true[-len(y_pred):], y_pred, *args, **kwargs)

    return _wrapper


def _second_level(values: np.ndarray, tags: Sequence) -> List[str]:
    if values.shape[1] == len(tags):
        return [_tag_name(t) for t in tags]
    return [str(i) for i in range(values.shape[1])]


def _isoformat(index: pd.DatetimeIndex) -> np.ndarray:
    """``[ts.isoformat() for ts in index]`` -- vectorised for the common case (naive or UTC, whole seconds), which is where the
    per-timestamp loop costs more than the GPU work of a 10 000-row request."""
    tz = index.tz
    if len(index) and (tz is None or str(tz) in ("UTC", "utc")):
        naive = index.tz_localize(None) if tz is not None else index
        secs = naive.values.astype("datetime64[s]")
        if (secs == naive.values).all():  # whole seconds (whatever the index's resolution)
            out = np.datetime_as_string(secs, unit="s")
            return (np.char.add(out, "+00:00") if tz is not None else out).astype(object)
    return np.array([ts.isoformat() for ts in index], dtype=object)


def _time_columns(index, n: int, frequency: Optional[timedelta]):
    """ISO ``start`` strings and ``end = start + frequency`` for a DatetimeIndex; None otherwise."""
    if isinstance(index, pd.DatetimeIndex):
        start = _isoformat(index)
        end = _isoformat(index + frequency) if frequency is not None else np.full(n, None, dtype=object)
        return start, end
    return np.full(n, None, dtype=object), np.full(n, None, dtype=object)


def make_base_dataframe(
    tags: Union[List, List[str]],
    model_input: np.ndarray,
    model_output: np.ndarray,
    target_tag_list: Optional[List] = None,
    index: Optional[Union[np.ndarray, pd.Index]] = None,
    frequency: Optional[timedelta] = None,
) -> pd.DataFrame:
    """
    MultiIndex-column frame ``start, end, model-input/<tag>..., model-output/<target>...``.  The model output
    sets the length: input and index are clipped to their last ``len(model_output)`` rows.  Second-level names
    are the tag names when the widths match, else ``"0".."k-1"``.
    """
    target_tag_list = target_tag_list if target_tag_list is not None else tags
    model_output = np.asarray(getattr(model_output, "values", model_output))
    n = len(model_output)
    model_input = np.asarray(getattr(model_input, "values", model_input))[-n:, :] if n else np.asarray(getattr(model_input, "values", model_input))[:0, :]
    idx = index[-n:] if index is not None else pd.RangeIndex(n)
    if n == 0 and index is not None:
        idx = index[:0]
    if not isinstance(idx, pd.Index):
        idx = pd.Index(idx)
    start, end = _time_columns(idx, n, frequency)
    columns = [("start", ""), ("end", "")]
    blocks = [pd.DataFrame({0: start, 1: end}, index=idx)]
    for name, values, names in (("model-input", model_input, tags), ("model-output", model_output, target_tag_list)):
        if values is None:
            continue
        columns += [(name, s) for s in _second_level(values, list(names))]
        blocks.append(pd.DataFrame(values, index=idx))
    frame = pd.concat(blocks, axis=1)
    frame.columns = pd.MultiIndex.from_tuples(columns)
    return frame
