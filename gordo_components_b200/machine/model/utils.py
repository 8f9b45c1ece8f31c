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


_ZERO, _SUFFIX_UTC = ord("0"), np.frombuffer("+00:00".encode("utf-32-le"), dtype=np.uint32)


def _iso_seconds(secs: np.ndarray, utc: bool) -> Optional[np.ndarray]:
    """
    ``YYYY-MM-DDTHH:MM:SS[+00:00]`` for int64 seconds since the epoch, as an object array of str, with integer arithmetic on
    whole arrays (days -> civil date after H. Hinnant's algorithm; ~6x faster than numpy's datetime_as_string, which in turn is
    ~10x faster than a Python loop).  None when a year falls outside 0001..9999 (the caller takes the per-element path).
    """
    n = len(secs)
    days, sod = np.divmod(secs, 86400)
    z = days + 719468
    era = np.floor_divide(z, 146097)
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    day = doy - (153 * mp + 2) // 5 + 1
    month = np.where(mp < 10, mp + 3, mp - 9)
    year = yoe + era * 400 + (month <= 2)
    if n and (year.min() < 1 or year.max() > 9999):
        return None
    hh, rem = np.divmod(sod, 3600)
    mi, ss = np.divmod(rem, 60)
    width = 25 if utc else 19
    buf = np.empty((n, width), dtype=np.uint32)
    for col, (value, digits) in zip((0, 5, 8, 11, 14, 17), ((year, 4), (month, 2), (day, 2), (hh, 2), (mi, 2), (ss, 2))):
        v = value.astype(np.int64)
        for d in range(digits - 1, -1, -1):
            v, r = np.divmod(v, 10)
            buf[:, col + d] = r + _ZERO
    buf[:, 4] = buf[:, 7] = ord("-")
    buf[:, 10] = ord("T")
    buf[:, 13] = buf[:, 16] = ord(":")
    if utc:
        buf[:, 19:] = _SUFFIX_UTC
    return buf.view(f"<U{width}").ravel().astype(object)


def _isoformat(index: pd.DatetimeIndex) -> np.ndarray:
    """``[ts.isoformat() for ts in index]`` -- vectorised for the common case (naive or UTC, whole seconds), which is where the
    per-timestamp loop costs more than the GPU work of a 10 000-row request."""
    tz = index.tz
    if len(index) and (tz is None or str(tz) in ("UTC", "utc")):
        naive = index.tz_localize(None) if tz is not None else index
        secs = naive.values.astype("datetime64[s]")
        if (secs == naive.values).all():  # whole seconds (whatever the index's resolution)
            out = _iso_seconds(secs.astype(np.int64), tz is not None)
            if out is not None:
                return out
    return np.array([ts.isoformat() for ts in index], dtype=object)


def _time_columns(index, n: int, frequency: Optional[timedelta]):
    """ISO ``start`` strings and ``end = start + frequency`` for a DatetimeIndex; None otherwise."""
    if isinstance(index, pd.DatetimeIndex):
        start = _isoformat(index)
        end = _isoformat(index + frequency) if frequency is not None else np.full(n, None, dtype=object)
        return start, end
    return np.full(n, None, dtype=object), np.full(n, None, dtype=object)


@functools.lru_cache(maxsize=512)
def _column_index(columns: tuple) -> pd.MultiIndex:
    # every request of a model has the same columns; building the MultiIndex costs more than the GPU work of a small request
    return pd.MultiIndex.from_tuples(columns)


def base_blocks(tags, model_input, model_output, target_tag_list=None, index=None, frequency=None):
    """
    The pieces of ``make_base_dataframe``: (row index, [start/end block, model-input block, model-output block], column tuples).
    Callers that append more column blocks (the anomaly frame) assemble everything with one ``frame_from_blocks``.
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
    return idx, blocks, columns


def frame_from_blocks(index, blocks, columns) -> pd.DataFrame:
    """One frame from column blocks (arrays or frames on ``index``) under two-level ``columns``; each block keeps its dtype."""
    frames = [b if isinstance(b, pd.DataFrame) else pd.DataFrame(b, index=index) for b in blocks]
    frame = pd.concat(frames, axis=1, ignore_index=True) if len(frames) > 1 else frames[0].copy()
    frame.columns = _column_index(tuple(columns)).copy()
    return frame


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
    return frame_from_blocks(*base_blocks(tags, model_input, model_output, target_tag_list, index, frequency))
