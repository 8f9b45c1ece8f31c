"""
Wire formats, the resident model store and the request checks of the serving path (no GPU needed).  Modelled on
tests/gordo/server/test_utils.py (frame <-> dict / parquet round trips, _verify_dataframe) and
tests/gordo/server/test_gordo_server.py / test_anomaly_view.py (status codes of malformed requests).
"""
import json
import os

import numpy as np
import pandas as pd
import pytest
from sklearn.preprocessing import MinMaxScaler

from gordo_components_b200 import serializer, server
from gordo_components_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector
from gordo_components_b200.machine.model.models import KerasAutoEncoder

TAGS = ["tag-0", "tag-1", "tag-2"]


def _frame(rows=6, cols=TAGS, seed=0):
    idx = pd.date_range("2016-01-01", periods=rows, freq="10min", tz="UTC")
    return pd.DataFrame(np.random.default_rng(seed).random((rows, len(cols))), columns=cols, index=idx)


def _anomaly_like_frame():
    X = _frame()
    columns = pd.MultiIndex.from_tuples([("start", ""), ("model-output", "tag-0"), ("model-output", "tag-1"), ("total-anomaly-scaled", "")])
    data = pd.DataFrame(np.arange(24.0).reshape(6, 4), columns=columns, index=X.index)
    data[("start", "")] = [t.isoformat() for t in X.index]
    return data


def test_dict_round_trips():
    """test_utils.py: multi-level frames, plain frames, integer indexes."""
    df = _anomaly_like_frame()
    as_dict = server.dataframe_to_dict(df)
    json.dumps(as_dict)  # string keys only
    assert set(as_dict) == {"start", "model-output", "total-anomaly-scaled"} and set(as_dict["model-output"]) == {"tag-0", "tag-1"}
    back = server.dataframe_from_dict(as_dict)
    assert back.index.equals(df.index)
    np.testing.assert_array_equal(back["model-output"].values, df["model-output"].values)
    np.testing.assert_array_equal(back["total-anomaly-scaled"].values.ravel(), df["total-anomaly-scaled"].values.ravel())
    assert df.index.dtype != object  # the input frame's index is left alone

    plain = _frame()
    again = server.dataframe_from_dict(json.loads(json.dumps(server.dataframe_to_dict(plain))))
    pd.testing.assert_frame_equal(again, plain, check_freq=False)
    numbered = pd.DataFrame({"a": [1.0, 2.0, 3.0]}, index=[2, 0, 1])
    back = server.dataframe_from_dict(json.loads(json.dumps(server.dataframe_to_dict(numbered))))
    assert list(back.index) == [0, 1, 2] and list(back["a"]) == [2.0, 3.0, 1.0]  # integer index restored and sorted


def test_parquet_round_trips():
    for df in (_frame(), _anomaly_like_frame()):
        buf = server.dataframe_into_parquet_bytes(df)
        assert isinstance(buf, bytes) and buf[:4] == b"PAR1"
        pd.testing.assert_frame_equal(server.dataframe_from_parquet_bytes(buf), df, check_freq=False)


def test_verify_dataframe():
    """test_utils.py::test_verify_dataframe: relabel, reorder / select, refuse."""
    unlabeled = pd.DataFrame(np.zeros((2, 3)))
    assert list(server.verify_dataframe(unlabeled, TAGS).columns) == TAGS and list(unlabeled.columns) == [0, 1, 2]
    shuffled = _frame(cols=["tag-2", "extra", "tag-0", "tag-1"])
    assert list(server.verify_dataframe(shuffled, TAGS).columns) == TAGS
    wrong = server.verify_dataframe(pd.DataFrame(np.zeros((2, 4))), TAGS)
    assert isinstance(wrong, server.Reply) and wrong.status == 400 and "Unexpected features" in wrong.body["message"]
    multi = server.verify_dataframe(_anomaly_like_frame(), TAGS)
    assert isinstance(multi, server.Reply) and multi.status == 400 and "multi-level" in multi.body["message"]


@pytest.fixture
def project(tmp_path):
    meta = {"name": "machine-1", "dataset": {"tag_list": [{"name": t, "asset": None} for t in TAGS], "resolution": "10min"}}
    serializer.dump(DiffBasedAnomalyDetector(base_estimator=KerasAutoEncoder(kind="feedforward_hourglass")), str(tmp_path / "machine-1"), metadata=meta)
    serializer.dump(MinMaxScaler().fit(np.random.random((5, 3))), str(tmp_path / "scaler-only"), metadata={"name": "scaler-only", "dataset": {"tag_list": TAGS, "target_tag_list": TAGS[:2]}})
    os.makedirs(tmp_path / "not-a-model")
    return str(tmp_path)


def test_model_store(project):
    store = server.ModelStore(project)
    assert store.names() == ["machine-1", "scaler-only"]
    assert store.model("machine-1") is store.model("machine-1")  # stays loaded
    assert store.tags("machine-1") == TAGS and store.target_tags("machine-1") == TAGS
    assert store.target_tags("scaler-only") == TAGS[:2] and store.frequency("scaler-only") is None
    assert store.frequency("machine-1") == pd.tseries.frequencies.to_offset("10min")
    with pytest.raises(FileNotFoundError):
        store.model("not-a-model")
    small = server.ModelStore(project, max_models=1)
    first = small.model("machine-1")
    small.model("scaler-only")
    assert small.model("machine-1") is not first  # evicted and loaded again


def test_request_checks_answer_like_the_reference(project):
    """Status codes and messages of utils.extract_X_y / anomaly._create_anomaly_response, before any model arithmetic."""
    store = server.ModelStore(project)
    X = server.dataframe_to_dict(_frame())
    missing = server.anomaly_prediction(store, "machine-1", json={"y": X})
    assert missing.status == 400 and missing.body == {"message": 'Cannot predict without "X"'} and missing.content_type == "application/json"
    assert server.anomaly_prediction(store, "machine-1", files={}).status == 400
    no_y = server.anomaly_prediction(store, "machine-1", json={"X": X})
    assert no_y.status == 400 and no_y.body == {"message": "Cannot perform anomaly without 'y' to compare against."}
    wide = server.anomaly_prediction(store, "machine-1", json={"X": server.dataframe_to_dict(_frame(cols=TAGS + ["more"])), "y": X})
    # a frame carrying all expected tags is reduced to them and reaches the model -- which has no thresholds yet: the detector's
    # AttributeError is answered like the reference does (anomaly.py:46-52)
    assert wide.status == 422
    bad = server.anomaly_prediction(store, "machine-1", json={"X": server.dataframe_to_dict(_frame(cols=["a", "b"])), "y": X})
    assert bad.status == 400 and "Unexpected features" in bad.body["message"]
    bad_y = server.anomaly_prediction(store, "machine-1", files={"X": server.dataframe_into_parquet_bytes(_frame()),
                                                                     "y": server.dataframe_into_parquet_bytes(_frame(cols=["a", "b"]))})
    assert bad_y.status == 400 and "Unexpected features" in bad_y.body["message"]
    not_detector = server.anomaly_prediction(store, "scaler-only", json={"X": X, "y": server.dataframe_to_dict(_frame(cols=TAGS[:2]))})
    assert not_detector.status == 422 and "Model is not an AnomalyDetector" in not_detector.body["message"]
    assert server.anomaly_prediction(store, "nope", json={"X": X, "y": X}).status == 404
    assert server.prediction(store, "nope", json={"X": X}).status == 404
    # a transformer-only model is served through transform (base.py / model_io.get_model_output)
    ok = server.prediction(store, "scaler-only", json={"X": X})
    assert ok.status == 200 and set(ok.body["data"]) == {"start", "end", "model-input", "model-output"} and float(ok.body["time-seconds"]) >= 0
    assert set(ok.body["data"]["model-input"]) == set(TAGS) and set(ok.body["data"]["model-output"]) == {"0", "1", "2"}  # 3 outputs, 2 target names: positions


class EchoDetector:
    """A stand-in with the detector's serving surface (picklable, no GPU): output = 2 * input, one smoothed block."""

    def predict(self, X):
        return np.asarray(getattr(X, "values", X), dtype=np.float32) * 2

    def anomaly(self, X, y, frequency=None):
        from gordo_components_b200.machine.model import utils as model_utils

        out = self.predict(X)
        frame = model_utils.make_base_dataframe(tags=list(X.columns), model_input=X.values, model_output=out, target_tag_list=list(y.columns),
                                                index=X.index, frequency=frequency)
        diff = np.abs(out - y.values).astype(np.float32)
        for j, tag in enumerate(y.columns):
            frame[("tag-anomaly-scaled", tag)] = diff[:, j]
            frame[("smooth-tag-anomaly-scaled", tag)] = diff[:, j] * 0.5
        frame[("total-anomaly-scaled", "")] = (diff ** 2).mean(axis=1)
        frame[("smooth-total-anomaly-scaled", "")] = frame[("total-anomaly-scaled", "")] * 0.5
        return frame


def test_views_json_and_parquet_with_a_stand_in_model(tmp_path):
    """The whole request path around the model: JSON and parquet in and out, dropped smooth columns, the plain prediction view."""
    meta = {"name": "echo", "dataset": {"tag_list": TAGS, "resolution": "10min"}}
    serializer.dump(EchoDetector(), str(tmp_path / "echo"), metadata=meta)
    store = server.ModelStore(str(tmp_path))
    X = _frame(rows=8)
    want = store.model("echo").anomaly(X, X, frequency=store.frequency("echo"))

    reply = server.anomaly_prediction(store, "echo", json={"X": server.dataframe_to_dict(X), "y": server.dataframe_to_dict(X)})
    assert reply.status == 200 and reply.content_type == "application/json"
    body = json.loads(json.dumps(reply.body))  # what actually travels
    assert float(body["time-seconds"]) >= 0 and not any(k.startswith("smooth-") for k in body["data"])
    got = server.dataframe_from_dict(body["data"])
    np.testing.assert_array_equal(got["model-output"].values, want["model-output"].values)
    np.testing.assert_array_equal(got["total-anomaly-scaled"].values.ravel(), want["total-anomaly-scaled"].values.ravel())
    assert list(got["end"].values.ravel()) == list(want["end"].values.ravel()) and got.index.equals(want.index)
    everything = server.anomaly_prediction(store, "echo", json={"X": server.dataframe_to_dict(X), "y": server.dataframe_to_dict(X)}, all_columns=True)
    assert {"smooth-tag-anomaly-scaled", "smooth-total-anomaly-scaled"} <= set(everything.body["data"])

    unlabelled = X.copy()
    unlabelled.columns = [str(i) for i in range(len(TAGS))]
    files = {"X": server.dataframe_into_parquet_bytes(unlabelled), "y": server.dataframe_into_parquet_bytes(X)}
    reply = server.anomaly_prediction(store, "echo", files=files, fmt="parquet")
    assert reply.status == 200 and reply.content_type == "application/octet-stream"
    frame = server.dataframe_from_parquet_bytes(reply.body)
    assert not any(c[0].startswith("smooth-") for c in frame.columns)
    pd.testing.assert_frame_equal(frame, want.drop(columns=[c for c in want.columns if c[0].startswith("smooth-")]), check_freq=False)

    reply = server.prediction(store, "echo", json={"X": server.dataframe_to_dict(X)})
    assert reply.status == 200 and set(reply.body["data"]) == {"start", "end", "model-input", "model-output"}
    out = server.dataframe_from_dict(json.loads(json.dumps(reply.body))["data"])
    np.testing.assert_array_equal(out["model-output"].values, X.values.astype(np.float32) * 2)
    np.testing.assert_array_equal(out["model-input"].values, X.values)
    as_parquet = server.prediction(store, "echo", files={"X": server.dataframe_into_parquet_bytes(X)}, fmt="parquet")
    assert server.dataframe_from_parquet_bytes(as_parquet.body)["model-output"].shape == (8, 3)


@pytest.mark.parametrize("rows,window,thresholds,index_kind", [(1, None, True, "utc"), (30, 5, True, "utc"), (12, 4, False, "naive"), (9, None, False, "range")])
def test_json_reply_from_blocks_equals_the_frame_route(tmp_path, rows, window, thresholds, index_kind):
    """
    This package's detectors answer JSON requests from their column blocks (no DataFrame); the reply must be what the frame route
    gives -- ``dataframe_to_dict(model.anomaly(...))`` minus the smooth columns unless ``all_columns``.  The GPU score is mocked.
    """
    T = 3
    idx = {"utc": pd.date_range("2019-01-01", periods=rows, freq="10min", tz="UTC"), "naive": pd.date_range("2019-01-01", periods=rows, freq="10min"),
           "range": pd.RangeIndex(rows)}[index_kind]
    rng = np.random.default_rng(rows)
    X = pd.DataFrame(rng.random((rows, T)), index=idx, columns=TAGS)
    det = DiffBasedAnomalyDetector(base_estimator=KerasAutoEncoder(kind="feedforward_hourglass"), require_thresholds=thresholds, window=window,
                                   smoothing_method="sma" if window else None)
    if thresholds:
        det.feature_thresholds_, det.aggregate_threshold_ = pd.Series(np.ones(T), index=TAGS), 0.5
    res = {"model-output": rng.random((rows, T)).astype(np.float32), "tag-anomaly-scaled": rng.random((rows, T)).astype(np.float32),
           "total-anomaly-scaled": rng.random(rows).astype(np.float32), "tag-anomaly-unscaled": rng.random((rows, T)).astype(np.float32),
           "total-anomaly-unscaled": rng.random(rows).astype(np.float32)}
    if thresholds:
        res.update({"anomaly-confidence": rng.random((rows, T)).astype(np.float32), "total-anomaly-confidence": rng.random(rows).astype(np.float32)})
    serializer.dump(det, str(tmp_path / "m"), metadata={"name": "m", "dataset": {"tag_list": TAGS, "resolution": "10min"}})
    store = server.ModelStore(str(tmp_path))
    model = store.model("m")
    model._score = lambda *a, **k: dict(res)
    model._smoothing = lambda metric: np.asarray(metric, dtype=np.float32) * 0.5
    if index_kind == "range":
        payload = {"X": X.to_dict(), "y": X.to_dict()}
        payload = json.loads(json.dumps({k: {c: {str(i): v for i, v in col.items()} for c, col in d.items()} for k, d in payload.items()}))
    else:
        payload = json.loads(json.dumps({"X": server.dataframe_to_dict(X), "y": server.dataframe_to_dict(X)}))
    Xr = server.dataframe_from_dict(payload["X"])
    frame = model.anomaly(Xr, Xr, frequency=store.frequency("m"))
    for all_columns in (False, True):
        reply = server.anomaly_prediction(store, "m", json=payload, all_columns=all_columns)
        assert reply.status == 200
        want = frame if all_columns else frame.drop(columns=[c for c in frame.columns if c[0] in server.DELETED_FROM_RESPONSE_COLUMNS])
        assert json.dumps(reply.body["data"]) == json.dumps(server.dataframe_to_dict(want))
        as_parquet = server.anomaly_prediction(store, "m", json=payload, all_columns=all_columns, fmt="parquet")
        pd.testing.assert_frame_equal(server.dataframe_from_parquet_bytes(as_parquet.body), want, check_freq=False)
    assert bool(window) == any(k.startswith("smooth-") for k in server.anomaly_prediction(store, "m", json=payload, all_columns=True).body["data"])


def test_overridden_anomaly_is_respected(tmp_path):
    """A subclass that changes ``anomaly`` is served through its own method, not through the column-block shortcut."""
    serializer.dump(_Doubling(base_estimator=KerasAutoEncoder(kind="feedforward_hourglass"), require_thresholds=False), str(tmp_path / "m"),
                    metadata={"name": "m", "dataset": {"tag_list": TAGS}})
    store = server.ModelStore(str(tmp_path))
    assert not server._frame_is_from_blocks(store.model("m")) and server._frame_is_from_blocks(DiffBasedAnomalyDetector())
    X = server.dataframe_to_dict(_frame())
    reply = server.anomaly_prediction(store, "m", json={"X": X, "y": X})
    assert reply.status == 200 and set(reply.body["data"]) == {"only"}


class _Doubling(DiffBasedAnomalyDetector):
    def anomaly(self, X, y, frequency=None):
        return pd.DataFrame(X.values[:, :1] * 2, index=X.index, columns=pd.MultiIndex.from_tuples([("only", "")]))


def test_requests_through_a_bucket_equal_direct_requests(tmp_path):
    """``bucket=``: the coalescer's score arrays take the same road into the reply as the detector's own (both mocked here: no GPU)."""
    T, rows = 3, 6
    rng = np.random.default_rng(4)
    res = {"model-output": rng.random((rows, T)).astype(np.float32), "tag-anomaly-scaled": rng.random((rows, T)).astype(np.float32),
           "total-anomaly-scaled": rng.random(rows).astype(np.float32), "tag-anomaly-unscaled": rng.random((rows, T)).astype(np.float32),
           "total-anomaly-unscaled": rng.random(rows).astype(np.float32), "anomaly-confidence": rng.random((rows, T)).astype(np.float32),
           "total-anomaly-confidence": rng.random(rows).astype(np.float32)}
    for name in ("in-bucket", "outside"):
        det = DiffBasedAnomalyDetector(base_estimator=KerasAutoEncoder(kind="feedforward_hourglass"))
        det.feature_thresholds_, det.aggregate_threshold_ = pd.Series(np.ones(T), index=TAGS), 0.5
        serializer.dump(det, str(tmp_path / name), metadata={"name": name, "dataset": {"tag_list": TAGS, "resolution": "10min"}})
    store = server.ModelStore(str(tmp_path))
    for name in store.names():
        store.model(name)._score = lambda *a, **k: dict(res)
    calls = []

    class FakeCoalescer:
        def anomaly(self, slot, X, y):
            calls.append((slot, X.shape, list(X.columns)))
            return dict(res)

    bucket = server.ResidentBucket.__new__(server.ResidentBucket)  # the constructor packs weights on the device; the serving logic is what is under test
    bucket.names, bucket.slot, bucket.coalescer = ["in-bucket"], {"in-bucket": 0}, FakeCoalescer()
    X = _frame(rows=rows)
    payload = json.loads(json.dumps({"X": server.dataframe_to_dict(X), "y": server.dataframe_to_dict(X)}))
    direct = server.anomaly_prediction(store, "in-bucket", json=payload)
    through = server.anomaly_prediction(store, "in-bucket", json=payload, bucket=bucket)
    assert calls == [(0, (rows, T), TAGS)] and through.status == 200
    assert json.dumps(through.body["data"]) == json.dumps(direct.body["data"])
    pq = server.anomaly_prediction(store, "in-bucket", json=payload, bucket=bucket, fmt="parquet")
    pd.testing.assert_frame_equal(server.dataframe_from_parquet_bytes(pq.body), server.dataframe_from_parquet_bytes(server.anomaly_prediction(store, "in-bucket", json=payload, fmt="parquet").body))
    assert len(calls) == 2
    server.anomaly_prediction(store, "outside", json=payload, bucket=bucket)  # not in the bucket: the model's own path
    assert len(calls) == 2
    # which models a bucket takes: this package's detector around a bare fitted auto-encoder, no smoothing window
    det = store.model("in-bucket")
    assert not server.ResidentBucket.eligible(det)  # the network has no weights yet
    det.base_estimator.kwargs.update({"n_features": T, "n_features_out": T})
    det.base_estimator._prepare_model()
    assert not server.ResidentBucket.eligible(det)  # the error scaler is not fitted
    det.scaler.fit(rng.random((8, T)))
    assert server.ResidentBucket.eligible(det)
    from sklearn.preprocessing import QuantileTransformer
    affine, det.scaler = det.scaler, QuantileTransformer(n_quantiles=4).fit(rng.random((8, T)))
    assert not server.ResidentBucket.eligible(det)  # not affine: this model keeps its per-request path, the rest of the store is still served
    det.scaler = affine
    det.window = 12
    assert not server.ResidentBucket.eligible(det)
    det.window = None
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import MinMaxScaler as MM
    assert not server.ResidentBucket.eligible(DiffBasedAnomalyDetector(base_estimator=Pipeline([("s", MM()), ("m", det.base_estimator)]), require_thresholds=False))
    assert not server.ResidentBucket.eligible(EchoDetector())
