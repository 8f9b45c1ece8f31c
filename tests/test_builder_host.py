"""
Host logic of the builder (no GPU needed): evaluation metrics from the five column moments against sklearn itself, split
and score metadata, which machines may take the batched path.  Modelled on tests/gordo/builder/test_builder.py:117-540.
"""
import numpy as np
import pandas as pd
import pytest
from sklearn import metrics as sk_metrics
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.model_selection import KFold, TimeSeriesSplit
from sklearn.preprocessing import MinMaxScaler

from gordo_components_b200 import builder
from gordo_components_b200.machine.model.utils import metric_wrapper


def _frame(rows=200, tags=4, seed=0, start="2020-01-01"):
    rng = np.random.default_rng(seed)
    idx = pd.date_range(start, periods=rows, freq="10min", tz="UTC")
    return pd.DataFrame(rng.random((rows, tags)).astype(np.float32), index=idx, columns=[f"TAG {i}" for i in range(tags)])


def numpy_moments(yhat, y):
    """What gb_cv_moments computes, in numpy (float64)."""
    yhat, y = yhat.astype(np.float64), y.astype(np.float64)
    e, c = yhat - y, y - y[0]
    return np.stack([e.sum(0), (e * e).sum(0), np.abs(e).sum(0), c.sum(0), (c * c).sum(0)])


@pytest.mark.parametrize("scaled", [False, True])
def test_scores_from_moments_match_sklearn(scaled):
    rng = np.random.default_rng(3)
    K, n, T = 3, 57, 6
    y_all = (rng.random((400, T)) * np.array([1, 10, 100, 0.01, 5, 1]) + np.array([0, 50, -20, 3, 1000, 0])).astype(np.float32)
    y_all[:, 5] = 7.0  # a constant tag: sklearn's zero-variance conventions apply
    scaler = MinMaxScaler().fit(y_all) if scaled else None
    folds_y = [y_all[100 * k : 100 * k + n] for k in range(K)]
    folds_p = [(fy + rng.normal(0, 0.05, fy.shape) * np.abs(fy).mean(0) + 0.01).astype(np.float32) for fy in folds_y]
    folds_p[1][:, 5] = 7.0  # fold 1 predicts the constant tag exactly -> 1.0; the others miss it -> 0.0
    mom = np.stack([numpy_moments(p, fy) for p, fy in zip(folds_p, folds_y)])
    got = builder.scores_from_moments(mom, n, scaler.scale_ if scaled else None)
    assert list(got) == list(builder.MOMENT_METRICS)
    for name, (per_tag, averaged) in got.items():
        func = getattr(sk_metrics, name)
        assert per_tag.shape == (K, T) and averaged.shape == (K,)
        for k in range(K):
            yt, yp = folds_y[k].astype(np.float64), folds_p[k].astype(np.float64)
            if scaled:
                yt, yp = scaler.transform(yt), scaler.transform(yp)
            np.testing.assert_allclose(averaged[k], func(yt, yp), rtol=1e-9, atol=1e-12, err_msg=name)
            np.testing.assert_allclose(per_tag[k], func(yt, yp, multioutput="raw_values"), rtol=1e-9, atol=1e-12, err_msg=name)
    assert got["r2_score"][0][1, 5] == 1.0 and got["r2_score"][0][0, 5] == 0.0
    with pytest.raises(ValueError):
        builder.scores_from_moments(mom, n, None, ["max_error"])


def test_scores_block_has_the_reference_keys():
    y = _frame(40, 3)
    mom = np.stack([numpy_moments(y.values[:10] + 0.1, y.values[:10]), numpy_moments(y.values[10:20] * 0.9, y.values[10:20])])
    block = builder.scores_block(builder.scores_from_moments(mom, 10), list(y.columns))
    scorers = builder.build_metrics_dict(builder.metrics_from_list(None), y)
    assert set(block) == set(scorers)  # '<metric>-<tag with dashes>' per tag + '<metric>'
    assert "r2-score-TAG-1" in block and "mean-squared-error" in block
    one = block["mean-absolute-error-TAG-0"]
    assert set(one) == {"fold-mean", "fold-std", "fold-max", "fold-min", "fold-1", "fold-2"}
    assert one["fold-1"] == pytest.approx(0.1, rel=1e-5) and one["fold-max"] == max(one["fold-1"], one["fold-2"])
    assert one["fold-mean"] == pytest.approx((one["fold-1"] + one["fold-2"]) / 2) and one["fold-std"] == pytest.approx(abs(one["fold-1"] - one["fold-2"]) / 2)


def test_metrics_from_list_and_metric_scorers():
    """test_builder.py:432-468, 471-540."""
    default = builder.metrics_from_list(None)
    assert [f.__name__ for f in default] == ["explained_variance_score", "r2_score", "mean_squared_error", "mean_absolute_error"]
    assert builder.metrics_from_list(["sklearn.metrics.r2_score", "max_error"]) == [sk_metrics.r2_score, sk_metrics.max_error]
    with pytest.raises(AttributeError):
        builder.metrics_from_list(["sklearn.metrics.no_such_metric"])

    y = _frame(60, 2)

    class Shifted(BaseEstimator, RegressorMixin):
        def predict(self, X):
            return np.asarray(X) + 0.5

    for scaler in (None, "sklearn.preprocessing.MinMaxScaler", MinMaxScaler()):
        scorers = builder.build_metrics_dict([sk_metrics.mean_squared_error], y, scaler=scaler)
        assert list(scorers) == ["mean-squared-error-TAG-0", "mean-squared-error-TAG-1", "mean-squared-error"]
        got = scorers["mean-squared-error-TAG-1"](Shifted(), y, y)
        rng_ = 1.0 if scaler is None else float(y["TAG 1"].max() - y["TAG 1"].min())
        assert got == pytest.approx((0.5 / rng_) ** 2, rel=1e-4)
    # the wrapper tail-aligns targets with shorter predictions (LSTM offset)
    assert metric_wrapper(sk_metrics.mean_absolute_error)(np.arange(10.0).reshape(-1, 1), np.arange(4.0, 10.0).reshape(-1, 1)) == 0.0


def test_build_split_dict():
    """test_builder.py:117-158."""
    X = _frame(20, 2)
    d = builder.build_split_dict(X, TimeSeriesSplit(n_splits=3))
    assert d["fold-1-n-train"] == 5 and d["fold-1-n-test"] == 5 and d["fold-3-n-train"] == 15
    assert d["fold-1-train-start"] == X.index[0] and d["fold-1-train-end"] == X.index[4]
    assert d["fold-3-test-start"] == X.index[15] and d["fold-3-test-end"] == X.index[19]
    assert len(d) == 18
    k = builder.build_split_dict(X, KFold(n_splits=2))
    assert k["fold-1-n-test"] == 10 and k["fold-2-train-start"] == X.index[0]


AE = {"gordo.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass", "epochs": 2}}
DETECTOR = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": AE}}


def _machine(name="m", model=DETECTOR, rows=200, **extra):
    X = _frame(rows, 4)
    return {"name": name, "model": model, "dataset": {"X": X, "y": X}, **extra}


def test_which_machines_take_the_batched_path():
    c = builder._canonical(0, _machine())
    assert c is not None and c.fit == {"epochs": 2, "batch_size": 32, "shuffle": True} and c.n_splits == 3
    assert c.spec.dims[0] == 4 and c.spec.dims[-1] == 4
    assert c.bucket() == builder._canonical(1, _machine("other")).bucket()
    assert c.bucket() != builder._canonical(1, _machine("longer", rows=300)).bucket()
    assert c.bucket() != builder._canonical(1, _machine("seeded", evaluation={"seed": 5})).bucket()
    five = builder._canonical(0, _machine(evaluation={"cv": {"sklearn.model_selection.TimeSeriesSplit": {"n_splits": 5}}, "metrics": ["r2_score"], "scoring_scaler": None}))
    assert five is not None and five.n_splits == 5

    # gordo's example config: the network behind one MinMaxScaler -- batched too, in its own bucket
    scaled = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {"sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler", AE]}}}}
    sc = builder._canonical(0, _machine(model=scaled))
    assert sc is not None and sc.input_scaler and not c.input_scaler and sc.bucket() != c.bucket() and sc.bucket()[:-1] == c.bucket()[:-1]
    pipeline = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {"sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.StandardScaler", AE]}}}}
    ranged = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {"sklearn.pipeline.Pipeline": {"steps": [
        {"sklearn.preprocessing.MinMaxScaler": {"feature_range": [-1, 1]}}, AE]}}}}
    three = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {"sklearn.pipeline.Pipeline": {"steps": [
        "gordo.machine.model.transformers.imputer.InfImputer", "sklearn.preprocessing.MinMaxScaler", AE]}}}}
    lstm = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {"gordo.machine.model.models.KerasLSTMAutoEncoder": {"kind": "lstm_hourglass", "lookback_window": 3}}}}
    smooth = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": AE, "window": 12}}
    robust = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": AE, "scaler": "sklearn.preprocessing.RobustScaler"}}
    kfcv = {"gordo.machine.model.anomaly.diff.DiffBasedKFCVAnomalyDetector": {"base_estimator": AE}}
    stopping = {"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {"gordo.machine.model.models.KerasAutoEncoder": {
        "kind": "feedforward_hourglass", "validation_split": 0.1, "callbacks": [{"tensorflow.keras.callbacks.EarlyStopping": {"patience": 1}}]}}}}
    for model in (pipeline, ranged, three, lstm, smooth, robust, kfcv, stopping, AE):
        assert builder._canonical(0, _machine(model=model)) is None
    for evaluation in ({"cv_mode": "cross_val_only"}, {"metrics": ["max_error"]}, {"scoring_scaler": "sklearn.preprocessing.StandardScaler"},
                       {"cv": {"sklearn.model_selection.KFold": {"n_splits": 3}}}, {"cv": {"sklearn.model_selection.TimeSeriesSplit": {"n_splits": 3, "gap": 2}}}):
        assert builder._canonical(0, _machine(evaluation=evaluation)) is None
    assert builder._canonical(0, _machine(rows=3)) is None


def test_machine_validation_and_datasets():
    with pytest.raises(ValueError):
        builder.ModelBuilder({"name": "x", "model": AE})
    with pytest.raises(ValueError):
        builder.FleetModelBuilder([_machine("a"), _machine("a")])
    with pytest.raises(TypeError):
        builder._get_data("not a dataset")

    class DS:
        def get_data(self):
            X = _frame(10, 2)
            return X, X

        def get_metadata(self):
            return {"rows": 10}

        def to_dict(self):
            return {"type": "DS"}

    X, y, meta = builder._get_data(DS())
    assert len(X) == 10 and meta == {"rows": 10}
    X, y, meta = builder._get_data((np.zeros((5, 2)), None))
    assert isinstance(X, pd.DataFrame) and y is X
    out = builder._machine_out({"name": "n", "model": AE, "dataset": DS(), "metadata": {"user_defined": {"k": 1}}}, {"model": {}, "dataset": {}})
    assert out["dataset"] == {"type": "DS"} and out["metadata"]["user_defined"] == {"k": 1} and out["metadata"]["build_metadata"] == {"model": {}, "dataset": {}}
    assert out["evaluation"]["cv_mode"] == "full_build"


def test_shards_cover_the_project_once():
    machines = [_machine(f"m-{i}", rows=40) for i in range(11)]
    full = builder.FleetModelBuilder(machines)
    for world in (1, 2, 4, 8, 16):
        names = [m["name"] for r in range(world) for m in full.shard(r, world).machines]
        assert names == [m["name"] for m in machines]  # contiguous blocks, in order, nothing twice
        sizes = [len(full.shard(r, world).machines) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_fleet_builder_control_flow(monkeypatch, tmp_path):
    """Bucketing, result order, per-machine fallback and the directory layout, with the two GPU entry points replaced."""
    calls = []

    def fake_bucket(members):
        calls.append([c.machine["name"] for c in members])
        if len(members[0].X) == 300:
            raise RuntimeError("does not fit")
        return [(f"batched:{c.machine['name']}", builder._machine_out(c.machine, {"model": {}, "dataset": {}})) for c in members]

    def fake_single(self, output_dir=None):
        calls.append(("single", self.machine["name"]))
        return f"single:{self.machine['name']}", builder._machine_out(self.machine, {"model": {}, "dataset": {}})

    monkeypatch.setattr(builder.FleetModelBuilder, "_build_bucket", staticmethod(fake_bucket))
    monkeypatch.setattr(builder.ModelBuilder, "build", fake_single)
    kfcv = {"gordo.machine.model.anomaly.diff.DiffBasedKFCVAnomalyDetector": {"base_estimator": AE}}
    machines = [_machine("a"), _machine("k", model=kfcv), _machine("big-1", rows=300), _machine("b"), _machine("big-2", rows=300)]
    results = builder.FleetModelBuilder(machines).build(str(tmp_path))
    assert [m for m, _ in results] == ["batched:a", "single:k", "single:big-1", "batched:b", "single:big-2"]
    assert calls == [("single", "k"), ["a", "b"], ["big-1", "big-2"], ("single", "big-1"), ("single", "big-2")]
    import os

    from gordo_components_b200 import serializer

    assert sorted(os.listdir(tmp_path)) == ["a", "b", "big-1", "big-2", "k"]
    assert serializer.load(str(tmp_path / "big-2")) == "single:big-2" and serializer.load_metadata(str(tmp_path / "a"))["name"] == "a"


PROJECT = """
machines:
  - name: pipe-1
    dataset: |
      tags: [TAG 1, TAG 2, TAG 3]
      target_tag_list: [TAG 3, TAG 4]
      train_start_date: '2019-01-01T00:00:00+00:00'
      train_end_date: '2019-01-03T00:00:00+00:00'
      data_provider:
        type: RandomDataProvider
    metadata: |
      information: first machine
    model: |
      sklearn.pipeline.Pipeline:
        steps:
          - sklearn.preprocessing.MinMaxScaler
          - sklearn.multioutput.MultiOutputRegressor:
              estimator: sklearn.linear_model.LinearRegression
  - name: pipe-2
    dataset:
      tag_list: [A, B]
      train_start_date: '2019-01-01T00:00:00+00:00'
      train_end_date: '2019-01-02T00:00:00+00:00'
      type: RandomDataset
    evaluation:
      cv_mode: cross_val_only
      metrics: [r2_score]
    runtime:
      builder: {resources: {limits: {memory: 1}}}
globals:
  model:
    sklearn.linear_model.Ridge:
      alpha: 0.5
  dataset:
    resolution: 30min
  evaluation:
    scoring_scaler: null
  metadata:
    owner: someone
  runtime:
    builder: {resources: {limits: {memory: 9, cpu: 2}}}
"""


def test_machines_from_config_and_local_build():
    """gordo/builder/local_build.py + Machine.from_config (machine.py:78-149): text blocks, globals, build of scikit-learn models."""
    machines = builder.machines_from_config(PROJECT, project_name="proj")
    one, two = machines
    assert [m["name"] for m in machines] == ["pipe-1", "pipe-2"] and one["project_name"] == "proj"
    assert "sklearn.pipeline.Pipeline" in one["model"] and two["model"] == {"sklearn.linear_model.Ridge": {"alpha": 0.5}}  # machine model wins, else globals
    assert one["evaluation"] == {**builder.DEFAULT_EVALUATION, "scoring_scaler": None}
    assert two["evaluation"]["cv_mode"] == "cross_val_only" and two["evaluation"]["metrics"] == ["r2_score"] and two["evaluation"]["scoring_scaler"] is None
    assert one["metadata"]["user_defined"] == {"global-metadata": {"owner": "someone"}, "machine-metadata": {"information": "first machine"}}
    assert two["runtime"] == {"builder": {"resources": {"limits": {"memory": 1, "cpu": 2}}}}  # globals patched by the machine
    X, y = one["dataset"].get_data()
    assert list(X.columns) == ["TAG 1", "TAG 2", "TAG 3"] and list(y.columns) == ["TAG 3", "TAG 4"]
    assert len(X) == 96 and X.index[0] == pd.Timestamp("2019-01-01T00:00:00+00:00") and (X.index[1] - X.index[0]) == pd.Timedelta("30min")  # resolution from globals
    np.testing.assert_array_equal(X["TAG 3"].values, y["TAG 3"].values)
    pd.testing.assert_frame_equal(one["dataset"].get_data()[0], X)  # seeded by the tag names
    assert one["dataset"].to_dict()["tag_list"] == ["TAG 1", "TAG 2", "TAG 3"] and one["dataset"].to_dict()["resolution"] == "30min"

    built = list(builder.local_build(PROJECT))
    assert [m["name"] for _, m in built] == ["pipe-1", "pipe-2"]
    (model1, m1), (model2, m2) = built
    scores = m1["metadata"]["build_metadata"]["model"]["cross_validation"]["scores"]
    assert "r2-score-TAG-4" in scores and "mean-absolute-error" in scores and len(scores) == 4 * 3
    assert m1["metadata"]["build_metadata"]["model"]["model_offset"] == 0 and model1.predict(X).shape == (96, 2)
    assert m1["dataset"]["tag_list"] == ["TAG 1", "TAG 2", "TAG 3"] and m1["metadata"]["build_metadata"]["dataset"]["dataset_meta"]["resolution"] == "30min"
    assert set(m2["metadata"]["build_metadata"]["model"]) == {"cross_validation"} and set(m2["metadata"]["build_metadata"]["model"]["cross_validation"]["scores"]) == {"r2-score-A", "r2-score-B", "r2-score"}
    assert [m["name"] for _, m in builder.local_build(PROJECT, batched=False)] == ["pipe-1", "pipe-2"]

    # real data comes in through datasets=; configs without a random provider say so
    frame = _frame(60, 2)
    custom = builder.machines_from_config({"machines": [{"name": "x", "model": AE, "dataset": {"tag_list": ["a"]}}]}, datasets={"x": (frame, frame)})
    assert custom[0]["dataset"][0] is frame
    called = builder.machines_from_config({"machines": [{"name": "x", "model": AE}]}, datasets=lambda m: ("made for", m["name"]))
    assert called[0]["dataset"] == ("made for", "x")
    with pytest.raises(TypeError):
        builder.machines_from_config({"machines": [{"name": "x", "model": AE, "dataset": {"tag_list": ["a"], "type": "TimeSeriesDataset"}}]})
    for bad in ({"machines": []}, {"machines": [{"model": AE}]}, {"machines": [{"name": "x"}]}):
        with pytest.raises(ValueError):
            builder.machines_from_config(bad)
    assert builder.patch_dict({"a": {"b": 1, "c": 2}}, {"a": {"b": 10}, "d": 4}) == {"a": {"b": 10, "c": 2}, "d": 4}
