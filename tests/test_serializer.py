"""
Definition <-> object graph and the model directory layout (no GPU needed).  Modelled on the reference's
tests/gordo/serializer/{test_serializer_from_definition,test_serializer_into_definition,test_serializer_load_dump}.py and
tests/gordo/machine/model/test_transformers.py.
"""
import os
import pickle

import numpy as np
import pandas as pd
import pytest
import yaml
from sklearn.compose import TransformedTargetRegressor
from sklearn.decomposition import PCA, TruncatedSVD
from sklearn.multioutput import MultiOutputRegressor
from sklearn.pipeline import FeatureUnion, Pipeline
from sklearn.preprocessing import FunctionTransformer, MinMaxScaler

from gordo_components_b200 import serializer
from gordo_components_b200.machine.model.anomaly.diff import DiffBasedAnomalyDetector, DiffBasedKFCVAnomalyDetector
from gordo_components_b200.machine.model.models import EarlyStopping, KerasAutoEncoder, KerasLSTMAutoEncoder
from gordo_components_b200.machine.model.register import register_model_builder
from gordo_components_b200.machine.model.transformer_funcs.general import multiply_by
from gordo_components_b200.machine.model.transformers.imputer import InfImputer


class DefinitionTestModel:
    """A class that builds itself: the ``from_definition`` / ``into_definition`` hooks (definition_test_model.py)."""

    @classmethod
    def from_definition(cls, definition: dict):
        return cls(int(definition.get("depth", 10)))

    def __init__(self, depth):
        self.depth = depth

    def into_definition(self):
        return {"depth": self.depth}


# ---------------------------------------------------------------- from_definition (test_serializer_from_definition.py:27-82)
MODELS_IN_PARAMS = [
    # a class path as a parameter value -> default instance
    {"sklearn.multioutput.MultiOutputRegressor": {"estimator": "sklearn.tree.DecisionTreeRegressor"}},
    # a one-key mapping as a parameter value -> instance with kwargs
    {"sklearn.multioutput.MultiOutputRegressor": {"estimator": {"sklearn.tree.DecisionTreeRegressor": {"max_depth": 3}}}},
    # a Pipeline as a parameter value, itself holding definitions
    {"sklearn.multioutput.MultiOutputRegressor": {"estimator": {"sklearn.pipeline.Pipeline": {"steps": [
        "sklearn.preprocessing.StandardScaler", {"sklearn.linear_model.Ridge": {"alpha": 0.5}}]}}}},
    # ... and a function path inside that Pipeline becomes the function
    {"sklearn.multioutput.MultiOutputRegressor": {"estimator": {"sklearn.pipeline.Pipeline": {"steps": [
        {"sklearn.cluster.FeatureAgglomeration": {"n_clusters": 3, "pooling_func": "numpy.median"}}, "sklearn.linear_model.LinearRegression"]}}}},
]


@pytest.mark.parametrize("definition", MODELS_IN_PARAMS)
def test_models_taking_models_as_parameters(definition):
    X, y = np.random.random((10, 10)), np.random.random((10, 2))
    model = serializer.from_definition(definition)
    assert isinstance(model, MultiOutputRegressor)
    model.fit(X, y)
    assert model.predict(X).shape == (10, 2)


def test_from_definition_hook_gets_raw_params():
    model = serializer.from_definition({f"{__name__}.DefinitionTestModel": {"depth": "300"}})
    assert type(model) is DefinitionTestModel and model.depth == 300
    assert type(serializer.from_definition(f"{__name__}.DefinitionTestModel")) is DefinitionTestModel
    assert serializer.into_definition(model) == {f"{__name__}.DefinitionTestModel": {"depth": 300}}


def test_definition_is_not_modified_and_unknown_paths_raise():
    definition = {"sklearn.pipeline.Pipeline": {"steps": ["sklearn.preprocessing.MinMaxScaler"]}}
    before = yaml.dump(definition)
    serializer.from_definition(definition)
    assert yaml.dump(definition) == before
    with pytest.raises(ImportError):
        serializer.from_definition({"sklearn.nothing.Here": {"a": 1}})
    with pytest.raises(ValueError):
        serializer.from_definition(3)
    with pytest.raises(ValueError):
        serializer.from_definition({"sklearn.pipeline.Pipeline": {"memory": None}})
    # plain strings stay plain strings
    assert serializer.locate("tanh") is None and serializer.locate("1.5") is None and serializer.locate("no.such.module") is None


def _all_kinds():
    for cls_name, kinds in register_model_builder.factories.items():
        for kind in kinds:
            yield cls_name, kind


# the same estimator graph twice: every keyword spelled out (as into_definition writes it), and the short list forms
FULL = """
sklearn.pipeline.Pipeline:
    memory:
    verbose: false
    steps:
        - sklearn.preprocessing._function_transformer.FunctionTransformer:
            func: gordo.machine.model.transformer_funcs.general.multiply_by
            kw_args:
                factor: 2
        - sklearn.pipeline.FeatureUnion:
            n_jobs: 1
            transformer_weights:
            transformer_list:
            - sklearn.pipeline.Pipeline:
                memory:
                steps:
                - sklearn.preprocessing.MinMaxScaler:
                    feature_range: [-1, 1]
                    copy: true
                - sklearn.decomposition.TruncatedSVD:
                    n_components: 2
            - sklearn.decomposition.PCA:
                n_components: 3
                whiten: false
                random_state:
        - sklearn.decomposition.PCA:
            n_components: 4
        - gordo.machine.model.models.{cls}:
            kind: {kind}
"""
SHORT = """
sklearn.pipeline.Pipeline:
    - sklearn.preprocessing.FunctionTransformer:
        func: gordo.machine.model.transformer_funcs.general.multiply_by
        kw_args: {{factor: 2}}
    - sklearn.pipeline.FeatureUnion:
        - sklearn.pipeline.Pipeline:
            - sklearn.preprocessing.MinMaxScaler:
                feature_range: [-1, 1]
            - sklearn.decomposition.TruncatedSVD:
                n_components: 2
        - sklearn.decomposition.PCA:
            n_components: 3
    - sklearn.decomposition.PCA:
        n_components: 4
    - gordo.machine.model.models.{cls}:
        kind: {kind}
"""


@pytest.mark.parametrize("template", [FULL, SHORT])
@pytest.mark.parametrize("cls_name,kind", list(_all_kinds()))
def test_pipeline_definitions_for_every_registered_kind(template, cls_name, kind):
    """test_serializer_from_definition.py:84-272: reference class paths resolve to this package's estimators."""
    pipe = serializer.from_definition(yaml.safe_load(template.format(cls=cls_name, kind=kind)))
    assert isinstance(pipe, Pipeline) and [name for name, _ in pipe.steps] == ["step_0", "step_1", "step_2", "step_3"]
    func, union, pca, model = (s for _, s in pipe.steps)
    assert isinstance(func, FunctionTransformer) and func.func is multiply_by and func.kw_args == {"factor": 2}
    assert isinstance(union, FeatureUnion) and [n for n, _ in union.transformer_list] == ["step_0", "step_1"]
    inner, inner_pca = (t for _, t in union.transformer_list)
    assert isinstance(inner, Pipeline) and isinstance(inner.steps[0][1], MinMaxScaler) and isinstance(inner.steps[1][1], TruncatedSVD)
    assert inner.steps[0][1].feature_range == (-1, 1)  # a YAML list became the tuple sklearn expects
    assert isinstance(inner_pca, PCA) and inner_pca.n_components == 3 and isinstance(pca, PCA) and pca.n_components == 4
    assert type(model).__name__ == cls_name and type(model).__module__ == "gordo_components_b200.machine.model.models"
    assert model.kind == kind
    # and back (test_serializer_into_definition.py:191-298)
    again = serializer.from_definition(serializer.into_definition(pipe))
    assert serializer.into_definition(again) == serializer.into_definition(pipe)


def test_detector_definitions():
    definition = yaml.safe_load(
        """
    gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector:
      scaler: sklearn.preprocessing.MinMaxScaler
      require_thresholds: false
      window: 144
      base_estimator:
        sklearn.compose.TransformedTargetRegressor:
          transformer: sklearn.preprocessing.MinMaxScaler
          regressor:
            sklearn.pipeline.Pipeline:
              steps:
                - sklearn.preprocessing.MinMaxScaler
                - gordo.machine.model.models.KerasAutoEncoder:
                    kind: feedforward_hourglass
                    batch_size: 3
                    compression_factor: 0.5
                    encoding_layers: 1
                    func: tanh
                    out_func: linear
                    epochs: 1
    """
    )
    det = serializer.from_definition(definition)
    assert type(det) is DiffBasedAnomalyDetector and det.window == 144 and det.require_thresholds is False
    assert isinstance(det.scaler, MinMaxScaler)
    ttr = det.base_estimator
    assert isinstance(ttr, TransformedTargetRegressor) and isinstance(ttr.transformer, MinMaxScaler)
    ae = ttr.regressor.steps[-1][1]
    assert type(ae) is KerasAutoEncoder and ae.kind == "feedforward_hourglass" and ae.kwargs["compression_factor"] == 0.5
    # into_definition(from_definition(x)) is the expansion `gordo build` uses as its cache key: it must be a fixed point
    expanded = serializer.into_definition(det)
    assert serializer.into_definition(serializer.from_definition(expanded)) == expanded
    assert yaml.safe_load(yaml.safe_dump(expanded)) == expanded  # primitives only

    kf = serializer.from_definition({"gordo.machine.model.anomaly.diff.DiffBasedKFCVAnomalyDetector": {
        "threshold_percentile": 0.9, "base_estimator": {"gordo.machine.model.models.KerasLSTMAutoEncoder": {"kind": "lstm_hourglass", "lookback_window": 4}}}})
    assert type(kf) is DiffBasedKFCVAnomalyDetector and kf.threshold_percentile == 0.9 and type(kf.base_estimator) is KerasLSTMAutoEncoder
    # a detector directly around a model with hooks: the detector forwards unknown attributes to it, yet its definition is its own
    plain = serializer.from_definition({"gordo.machine.model.anomaly.diff.DiffBasedAnomalyDetector": {"base_estimator": {"gordo.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass", "epochs": 3}}}})
    d = serializer.into_definition(plain)["gordo_components_b200.machine.model.anomaly.diff.DiffBasedAnomalyDetector"]
    assert d["base_estimator"] == {"gordo_components_b200.machine.model.models.KerasAutoEncoder": {"kind": "feedforward_hourglass", "epochs": 3}} and d["shuffle"] is False
    assert serializer.into_definition(serializer.from_definition(serializer.into_definition(plain))) == serializer.into_definition(plain)
    # the pre-1.0 package layout still names the same classes
    old = serializer.from_definition({"gordo_components.model.models.KerasAutoEncoder": {"kind": "feedforward_symmetric"}})
    assert type(old) is KerasAutoEncoder


def test_callbacks_and_kwargs_are_kept(tmp_path):
    definition = yaml.safe_load(
        """
    gordo.machine.model.models.KerasAutoEncoder:
      kind: feedforward_hourglass
      epochs: 50
      validation_split: 0.2
      some_made_up_kwarg: 7
      callbacks:
        - tensorflow.keras.callbacks.EarlyStopping:
            monitor: val_loss
            patience: 3
    """
    )
    model = serializer.from_definition(definition)
    # test_captures_kwarg_to_init (test_serializer_into_definition.py:300-316): unknown kwargs ride along
    assert model.kwargs["some_made_up_kwarg"] == 7 and model.kwargs["callbacks"] == definition[next(iter(definition))]["callbacks"]
    assert serializer.into_definition(model) == {"gordo_components_b200.machine.model.models.KerasAutoEncoder": {**definition[next(iter(definition))]}}
    # as a plain parameter value the list is built (from_definition.py:317-318, 337-372)
    built = serializer.load_params_from_definition({"callbacks": definition[next(iter(definition))]["callbacks"], "epochs": 2})
    assert isinstance(built["callbacks"][0], EarlyStopping) and built["callbacks"][0].patience == 3 and built["epochs"] == 2
    with pytest.raises(ValueError):
        serializer.load_params_from_definition(["not", "a", "dict"])


def test_prune_default_params():
    pipe = Pipeline([("pca", PCA(n_components=4)), ("mm", MinMaxScaler(feature_range=(-1, 1)))])
    full = serializer.into_definition(pipe)["sklearn.pipeline.Pipeline"]
    assert set(full) >= {"steps", "memory", "verbose"}
    assert full["steps"][1] == {"sklearn.preprocessing._data.MinMaxScaler": {"feature_range": [-1, 1], "copy": True, "clip": False}}
    pruned = serializer.into_definition(pipe, prune_default_params=True)["sklearn.pipeline.Pipeline"]
    assert "memory" not in pruned and "verbose" not in pruned and len(pruned["steps"]) == 2
    rebuilt = serializer.from_definition(serializer.into_definition(pipe))
    assert rebuilt.steps[0][1].n_components == 4 and rebuilt.steps[1][1].feature_range == (-1, 1)


# ---------------------------------------------------------------- dump / load (test_serializer_load_dump.py:26-140)
@pytest.mark.parametrize("model", [
    lambda: KerasAutoEncoder(kind="feedforward_hourglass"),
    lambda: Pipeline([("mm", MinMaxScaler()), ("ae", KerasLSTMAutoEncoder(kind="lstm_symmetric", lookback_window=3))]),
    lambda: DiffBasedAnomalyDetector(base_estimator=KerasAutoEncoder(kind="feedforward_symmetric", dims=(4, 2), funcs=("tanh", "tanh"))),
])
def test_dump_load_models(tmp_path, model):
    model = model()
    dest = tmp_path / "deep" / "dir"
    serializer.dump(model, str(dest), metadata={"name": "m1", "n": 1}, info={"checksum": "abc"})
    assert sorted(os.listdir(dest)) == ["info.json", "metadata.json", "model.pkl"]
    loaded = serializer.load(str(dest))
    assert type(loaded) is type(model) and serializer.into_definition(loaded) == serializer.into_definition(model)
    assert serializer.load_metadata(str(dest)) == {"name": "m1", "n": 1}
    assert serializer.load_info(str(dest)) == {"checksum": "abc"}
    assert type(serializer.loads(serializer.dumps(model))) is type(model)
    assert pickle.loads(pickle.dumps(model)).get_params().keys() == model.get_params().keys()


@pytest.mark.parametrize("location", ("metadata.json", "../metadata.json", None))
def test_load_metadata_locations(tmp_path, location):
    model_dir = tmp_path / "project" / "model"
    os.makedirs(model_dir)
    if location is None:
        with pytest.raises(FileNotFoundError):
            serializer.load_metadata(str(model_dir))
        assert serializer.metadata_path(str(model_dir)) is None
        return
    with open(os.path.join(model_dir, location), "w") as f:
        f.write('{"key": "value"}')
    assert serializer.load_metadata(str(model_dir)) == {"key": "value"}
    assert os.path.samefile(serializer.metadata_path(str(model_dir)), os.path.join(model_dir, location))


# ---------------------------------------------------------------- transformers (test_transformers.py:17-175)
def test_multiply_by_in_function_transformer():
    X = np.random.random((10, 3))
    tf = FunctionTransformer(func=multiply_by, kw_args={"factor": 2})
    np.testing.assert_array_equal(tf.fit_transform(X), X * 2)
    with pytest.raises(TypeError):
        FunctionTransformer(func=multiply_by).fit_transform(X)  # factor is required


@pytest.mark.parametrize("strategy", ["extremes", "minmax"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_infimputer_strategies(strategy, dtype):
    rng = np.random.default_rng(0)
    base = rng.random((100, 10)).astype(dtype)
    flat = base.ravel()
    flat[rng.integers(0, flat.size, 100)] = np.inf
    flat[rng.integers(0, flat.size, 100)] = -np.inf
    pos, neg = np.isposinf(base), np.isneginf(base)
    assert pos.any() and neg.any()
    X = InfImputer(strategy=strategy, delta=2.0).fit_transform(base.copy())
    assert np.isfinite(X).all()
    np.testing.assert_array_equal(X[~(pos | neg)], base[~(pos | neg)])
    if strategy == "extremes":
        assert (X[pos] == np.finfo(dtype).max).all() and (X[neg] == np.finfo(dtype).min).all()
    else:
        masked = np.ma.masked_invalid(base)
        hi, lo = masked.max(axis=0).filled(np.nan), masked.min(axis=0).filled(np.nan)
        np.testing.assert_allclose(X[pos], (np.broadcast_to(hi, base.shape) + 2.0)[pos], rtol=1e-6)
        np.testing.assert_allclose(X[neg], (np.broadcast_to(lo, base.shape) - 2.0)[neg], rtol=1e-6)
    frame = InfImputer(strategy=strategy).fit_transform(pd.DataFrame(base.copy()))
    assert np.isfinite(np.asarray(frame)).all()


def test_infimputer_fill_values_and_definition():
    base = np.random.default_rng(1).random((100, 10)).astype(np.float32)
    base.ravel()[[1, 2, 3, 4, 5]] = np.inf
    base.ravel()[[6, 7, 8, 9, 10]] = -np.inf
    X = InfImputer(inf_fill_value=9999.0, neg_inf_fill_value=-9999.0).fit_transform(base.copy())
    assert (X.ravel()[[1, 2, 3, 4, 5]] == 9999.0).all() and (X.ravel()[[6, 7, 8, 9, 10]] == -9999.0).all()
    only_pos = InfImputer(inf_fill_value=5.0, strategy=None).fit_transform(base.copy())
    assert (only_pos.ravel()[[1, 2, 3, 4, 5]] == 5.0).all() and np.isneginf(only_pos.ravel()[[6, 7, 8, 9, 10]]).all()
    # near the dtype's edge the fill saturates instead of overflowing
    edge = np.array([[np.finfo(np.float32).max, 1.0], [np.inf, -np.inf]], dtype=np.float32)
    out = InfImputer(delta=2.0).fit_transform(edge.copy())
    assert out[1, 0] == np.finfo(np.float32).max and out[1, 1] == -1.0
    for text in ("sklearn.pipeline.Pipeline:\n  steps:\n    - gordo.machine.model.transformers.imputer.InfImputer",
                 "sklearn.pipeline.Pipeline:\n  steps:\n    - gordo.machine.model.transformers.imputer.InfImputer:\n        strategy: extremes\n        delta: 3.0"):
        pipe = serializer.from_definition(yaml.safe_load(text))
        assert isinstance(pipe.steps[0][1], InfImputer)
        assert serializer.from_definition(serializer.into_definition(pipe)).steps[0][1].get_params() == pipe.steps[0][1].get_params()
