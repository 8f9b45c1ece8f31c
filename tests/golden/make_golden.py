"""
Regenerates the golden fixtures in this directory by running the **reference's own code**
(unmodified, from /root/reference, via oracle/reference_loader.py).  Run in the build
container only -- /root/reference does not exist on the GPU box:

    python tests/golden/make_golden.py

Fixtures written:
  hourglass_dims.json        reference hourglass_calc_dims over a grid + the reference test table
  anomaly_<case>.npz         X, y, per-fold predictions/scalers/thresholds, and every column
                             block of DiffBasedAnomalyDetector.anomaly() from the reference
  ffnet_anomaly.npz          the same, with the base estimator being a fixed-weight hourglass
                             net (oracle/keras_math.ff_forward): pins net -> anomaly end to end
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import keras_math as km  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402

warnings.filterwarnings("ignore")
ref = load_reference()

from sklearn.base import BaseEstimator  # noqa: E402
from sklearn.linear_model import LinearRegression  # noqa: E402
from sklearn.model_selection import TimeSeriesSplit  # noqa: E402
from sklearn.multioutput import MultiOutputRegressor  # noqa: E402
from sklearn.preprocessing import MinMaxScaler  # noqa: E402


class FixedNet(BaseEstimator):
    """sklearn-style estimator around a fixed-weight oracle Dense stack (fit is a no-op)."""

    def __init__(self, n_features=8, seed=0):
        self.n_features = n_features
        self.seed = seed

    def _net(self):
        spec = km.ff_hourglass_spec(self.n_features)
        w = km.init_ff_weights(spec, np.random.default_rng(self.seed))
        # non-zero biases so the bias path is pinned too
        rng = np.random.default_rng(self.seed + 1)
        w = [(W, rng.uniform(-0.1, 0.1, size=b.shape).astype(np.float32)) for W, b in w]
        return spec, w

    def fit(self, X, y=None):
        return self

    def predict(self, X):
        spec, w = self._net()
        return km.ff_predict(spec, w, np.asarray(getattr(X, "values", X)))

    def score(self, X, y, sample_weight=None):
        return 0.0


def dims_fixture():
    table = []
    for cf in (0.0, 0.1, 0.2, 0.3, 0.5, 0.6, 0.75, 1.0):
        for layers in (1, 2, 3, 4, 5):
            for n in (1, 3, 4, 5, 8, 10, 64, 100, 128, 1000):
                table.append([cf, layers, n, list(ref.hourglass_calc_dims(cf, layers, n))])
    ref_test_table = [  # tests/gordo/machine/model/test_factories_utils.py:8-24
        [0.2, 4, 5, [4, 3, 2, 1]],
        [0.5, 3, 10, [8, 7, 5]],
        [0.5, 3, 3, [3, 2, 2]],
        [0.3, 3, 10, [8, 5, 3]],
        [1, 3, 10, [10, 10, 10]],
        [0, 3, 100000, [66667, 33334, 1]],
    ]
    for cf, layers, n, want in ref_test_table:
        assert list(ref.hourglass_calc_dims(cf, layers, n)) == want
    with open(os.path.join(HERE, "hourglass_dims.json"), "w") as f:
        json.dump({"grid": table, "reference_test_table": ref_test_table}, f)


def anomaly_fixture(name, n_rows, n_tags, window, method, datetime_index, base="linear", seed=0):
    rng = np.random.default_rng(seed)
    cols = [f"tag-{i}" for i in range(n_tags)]
    index = pd.date_range("2019-01-01", periods=n_rows, freq="10min", tz="UTC") if datetime_index else pd.RangeIndex(n_rows)
    X = pd.DataFrame(rng.random((n_rows, n_tags)), columns=cols, index=index)
    if base == "linear":
        y = pd.DataFrame(rng.random((n_rows, n_tags)) * np.arange(1, n_tags + 1), columns=cols, index=index)
        est = MultiOutputRegressor(LinearRegression())
    else:
        y = X.copy()
        est = FixedNet(n_features=n_tags, seed=seed)
    det = ref.DiffBasedAnomalyDetector(base_estimator=est, scaler=MinMaxScaler(), window=window, smoothing_method=method)
    cv = TimeSeriesSplit(n_splits=3)
    cvo = det.cross_validate(X=X, y=y, cv=cv)
    save = dict(X=X.values, y=y.values, n_splits=3, window=-1 if window is None else window, method=str(method),
                datetime_index=bool(datetime_index))
    for i, ((tr, te), fold) in enumerate(zip(cv.split(X, y), cvo["estimator"])):
        save[f"fold{i}_pred"] = np.asarray(fold.predict(X.iloc[te]))
        save[f"fold{i}_scale"] = fold.scaler.scale_
        save[f"fold{i}_min"] = fold.scaler.min_
        save[f"fold{i}_test_start"] = te[0]
        save[f"fold{i}_test_len"] = len(te)
    save["feature_thresholds_per_fold"] = det.feature_thresholds_per_fold_.values.astype(np.float64)
    save["aggregate_thresholds_per_fold"] = np.array([det.aggregate_thresholds_per_fold_[f"fold-{i}"] for i in range(3)])
    save["feature_thresholds"] = det.feature_thresholds_.values.astype(np.float64)
    save["aggregate_threshold"] = np.float64(det.aggregate_threshold_)
    if window is not None:
        save["smooth_feature_thresholds"] = det.smooth_feature_thresholds_.values.astype(np.float64)
        save["smooth_aggregate_threshold"] = np.float64(det.smooth_aggregate_threshold_)
    det.fit(X, y)
    save["scale"], save["min"] = det.scaler.scale_, det.scaler.min_
    save["pred"] = np.asarray(det.predict(X))
    frame = det.anomaly(X, y, frequency=pd.Timedelta("10min") if datetime_index else None)
    save["columns_level0"] = np.array(list(dict.fromkeys(frame.columns.get_level_values(0))))
    save["columns"] = np.array(["|".join(map(str, c)) for c in frame.columns])
    for top in dict.fromkeys(frame.columns.get_level_values(0)):
        block = frame[top]
        if top in ("start", "end"):
            save[f"frame_{top}"] = np.array([str(v) for v in np.asarray(block).ravel()])
        else:
            save[f"frame_{top}"] = np.asarray(block, dtype=np.float64)
    if base != "linear":
        spec, w = est._net()
        save["net_dims"] = np.array(spec.dims)
        for l, (W, b) in enumerate(w):
            save[f"W{l}"], save[f"b{l}"] = W, b
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **save)
    print(name, "ok", frame.shape)


def kfcv_fixture(name, n_rows, n_tags, window, method, q, seed):
    """DiffBasedKFCVAnomalyDetector (diff.py:461-635) from the reference itself, LinearRegression base estimator."""
    rng = np.random.default_rng(seed)
    cols = [f"tag-{i}" for i in range(n_tags)]
    index = pd.date_range("2019-01-01", periods=n_rows, freq="10min", tz="UTC")
    X = pd.DataFrame(rng.random((n_rows, n_tags)), columns=cols, index=index)
    y = pd.DataFrame(rng.random((n_rows, n_tags)) * np.arange(1, n_tags + 1), columns=cols, index=index)
    det = ref.DiffBasedKFCVAnomalyDetector(base_estimator=MultiOutputRegressor(LinearRegression()), scaler=MinMaxScaler(), window=window,
                                           smoothing_method=method, threshold_percentile=q)
    det.cross_validate(X=X, y=y)
    det.fit(X, y)
    frame = det.anomaly(X, y, frequency=pd.Timedelta("10min"))
    save = dict(X=X.values, y=y.values, window=window, method=str(method), q=q, feature_thresholds=np.asarray(det.feature_thresholds_, dtype=np.float64),
                aggregate_threshold=np.float64(det.aggregate_threshold_), columns_level0=np.array(list(dict.fromkeys(frame.columns.get_level_values(0)))))
    for top in ("total-anomaly-confidence", "anomaly-confidence", "smooth-total-anomaly-scaled", "smooth-tag-anomaly-unscaled"):
        save[f"frame_{top}"] = np.asarray(frame[top], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **save)
    print(name, "ok", frame.shape)


if __name__ == "__main__":
    dims_fixture()
    kfcv_fixture("kfcv_smm", 300, 3, 12, "smm", 0.99, seed=6)
    kfcv_fixture("kfcv_ewma", 400, 4, 24, "ewma", 0.9, seed=7)
    anomaly_fixture("anomaly_plain", 300, 3, None, None, False)
    anomaly_fixture("anomaly_smm", 300, 3, 12, "smm", True, seed=1)
    anomaly_fixture("anomaly_sma", 200, 4, 12, "sma", True, seed=2)
    anomaly_fixture("anomaly_ewma", 200, 4, 12, "ewma", False, seed=3)
    anomaly_fixture("ffnet_anomaly", 400, 8, None, None, True, base="net", seed=4)
    anomaly_fixture("ffnet_anomaly_t64", 200, 64, None, None, True, base="net", seed=5)
