"""
The builder side of the hot path: what ``gordo build`` does for one machine (``ModelBuilder``: data -> model from its
definition -> cross validation with the evaluation metrics -> final fit -> offset + metadata -> ``model.pkl`` /
``metadata.json``; gordo/builder/build_model.py:48-340, 345-570) and the same for a whole project at once
(``FleetModelBuilder``).

``FleetModelBuilder`` is where the batched kernels pay off: machines whose definition is the canonical
``DiffBasedAnomalyDetector(base_estimator=KerasAutoEncoder(<feed-forward kind>), scaler=MinMaxScaler())`` -- the network bare or
behind one ``MinMaxScaler`` in a Pipeline, as in gordo's example configs -- are bucketed by architecture and training length, and every bucket is built by ``fleet.build_fleet`` -- all final fits and all CV folds
in one ``gb_ffae_fit`` launch, fold scoring / thresholds / scaler statistics / metric moments one launch each.  The
cross-validation ``scores`` block of the metadata is then assembled on the host from ``gb_cv_moments``' five sums per
(fold, tag).  Any other definition (other transformers in a Pipeline, LSTM models, K-fold detectors, custom metrics ...) goes
through ``ModelBuilder``: one machine at a time, still on the GPU through the estimators' own fit / predict.

Machines are plain dicts in the layout of ``Machine.to_dict()`` (gordo/machine/machine.py:226-246): ``name``, ``model`` (a
definition), ``dataset``, and optionally ``project_name``, ``evaluation``, ``metadata``, ``runtime``.  ``dataset`` is
anything with ``get_data() -> (X, y)`` (and optionally ``get_metadata()``), an ``(X, y)`` pair, or ``{"X": ..., "y": ...}`` --
gordo's data providers themselves are outside this path.  Out of scope as well: the model cache / registry arguments of
``ModelBuilder.build`` and the reporters.
"""
import copy
import datetime
import importlib
import logging
import os
import random
import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd
from sklearn import metrics as sk_metrics
from sklearn.base import BaseEstimator
from sklearn.model_selection import TimeSeriesSplit, cross_validate
from sklearn.pipeline import Pipeline
from sklearn.preprocessing import MinMaxScaler

from . import __version__, serializer
from .machine.model.base import GordoBase
from .machine.model.utils import metric_wrapper

logger = logging.getLogger(__name__)

# NormalizedConfig.DEFAULT_CONFIG_GLOBALS["evaluation"] (gordo/workflow/config_elements/normalized_config.py:97-106)
DEFAULT_EVALUATION: Dict[str, Any] = {
    "cv_mode": "full_build",
    "scoring_scaler": "sklearn.preprocessing.MinMaxScaler",
    "metrics": ["explained_variance_score", "r2_score", "mean_squared_error", "mean_absolute_error"],
}
DEFAULT_CV = {"sklearn.model_selection.TimeSeriesSplit": {"n_splits": 3}}
MOMENT_METRICS = ("explained_variance_score", "r2_score", "mean_squared_error", "mean_absolute_error")


# ------------------------------------------------------------------------------------------------ pieces of ModelBuilder
def metrics_from_list(metric_list: Optional[Sequence[str]] = None) -> List[Callable]:
    """Metric names (looked up in ``sklearn.metrics``) or dotted function paths -> functions (build_model.py:671-707)."""
    funcs = []
    for path in metric_list or DEFAULT_EVALUATION["metrics"]:
        func = None
        if "." in path:
            module, _, name = path.rpartition(".")
            try:
                func = getattr(importlib.import_module(module), name, None)
            except ImportError:
                func = None
        if func is None:
            func = getattr(sk_metrics, path, None)
        if func is None:
            raise AttributeError(f"Could not locate metric function: {path}")
        funcs.append(func)
    return funcs


def _column_label(col) -> str:
    return str(col).replace(" ", "-")


def build_metrics_dict(metrics_list: Sequence[Callable], y: pd.DataFrame, scaler=None) -> dict:
    """
    sklearn scorers keyed ``<metric>-<tag>`` for every target tag and ``<metric>`` for the average over tags
    (build_model.py:378-446).  ``scaler`` (object or definition) is fitted on ``y`` and applied to targets and predictions
    before scoring.
    """
    if scaler:
        if isinstance(scaler, (str, dict)):
            scaler = serializer.from_definition(scaler)
        scaler.fit(y)

    def column_metric(func, index):
        def score(y_true, y_pred):
            y_true = getattr(y_true, "values", y_true)
            y_pred = getattr(y_pred, "values", y_pred)
            return func(y_true[:, index], y_pred[:, index])

        return score

    out = {}
    for func in metrics_list:
        name = func.__name__.replace("_", "-")
        for index, col in enumerate(y.columns):
            out[f"{name}-{_column_label(col)}"] = sk_metrics.make_scorer(metric_wrapper(column_metric(func, index), scaler=scaler))
        out[name] = sk_metrics.make_scorer(metric_wrapper(func, scaler=scaler))
    return out


def build_split_dict(X: pd.DataFrame, split_obj) -> dict:
    """Start / end timestamps and sizes of every CV fold's train and test part (build_model.py:347-376)."""
    out: Dict[str, Any] = {}
    for i, (train, test) in enumerate(split_obj.split(X), start=1):
        out[f"fold-{i}-train-start"] = X.index[train[0]]
        out[f"fold-{i}-train-end"] = X.index[train[-1]]
        out[f"fold-{i}-test-start"] = X.index[test[0]]
        out[f"fold-{i}-test-end"] = X.index[test[-1]]
        out[f"fold-{i}-n-train"] = len(train)
        out[f"fold-{i}-n-test"] = len(test)
    return out


def fold_summary(values) -> dict:
    """``fold-mean/std/max/min`` and ``fold-<i>`` of one metric's per-fold values (build_model.py:274-289)."""
    v = np.asarray(values, dtype=np.float64)
    out = {"fold-mean": float(v.mean()), "fold-std": float(v.std()), "fold-max": float(v.max()), "fold-min": float(v.min())}
    out.update({f"fold-{i + 1}": float(x) for i, x in enumerate(v)})
    return out


def determine_offset(model, X) -> int:
    """Rows the model's output is shorter than its input (LSTM look-back; build_model.py:448-471)."""
    X = getattr(X, "values", X)
    out = model.predict(X) if hasattr(model, "predict") else model.transform(X)
    return len(X) - len(out)


def extract_metadata_from_model(model, metadata: Optional[dict] = None) -> dict:
    """``get_metadata()`` of every GordoBase found in ``model`` (last Pipeline steps, estimator attributes; build_model.py:515-569)."""
    out = dict(metadata or {})
    if isinstance(model, Pipeline):
        out.update(extract_metadata_from_model(model.steps[-1][1]))
        return out
    if isinstance(model, GordoBase):
        out.update(model.get_metadata())
    for key, val in vars(model).items():
        if key == "regressor":  # TransformedTargetRegressor keeps the unfitted original next to regressor_
            continue
        if isinstance(val, Pipeline):
            out.update(extract_metadata_from_model(val.steps[-1][1]))
        elif isinstance(val, (GordoBase, BaseEstimator)):
            out.update(extract_metadata_from_model(val))
    return out


def scores_from_moments(moments: np.ndarray, n_rows: int, scoring_scale: Optional[np.ndarray] = None,
                        metrics: Sequence[str] = MOMENT_METRICS) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """
    The four evaluation metrics of one machine from ``gb_cv_moments``: ``moments`` is ``[folds, 5, tags]`` (sum e, sum e^2,
    sum |e|, sum (y-y0), sum (y-y0)^2 over the fold's ``n_rows`` test rows, e = prediction - target), ``scoring_scale`` the
    per-tag ``scale_`` of the scoring scaler fitted on all targets (an affine map per tag: errors scale by it, the two
    ratio metrics do not change).  Returns ``{metric: (per_tag [folds, tags], averaged [folds])}`` with sklearn's
    conventions: uniform average over tags; a constant target scores 1 when predicted exactly and 0 otherwise.
    """
    m = np.asarray(moments, dtype=np.float64)
    n = float(n_rows)
    se, see, sae, sc, scc = (m[..., q, :] for q in range(5))
    s = 1.0 if scoring_scale is None else np.asarray(scoring_scale, dtype=np.float64)

    def explained(numerator, denominator):
        out = np.ones_like(numerator)
        ok = (numerator != 0) & (denominator != 0)
        out[ok] = 1.0 - numerator[ok] / denominator[ok]
        out[(numerator != 0) & (denominator == 0)] = 0.0
        return out

    tss = np.maximum(scc - sc * sc / n, 0.0)  # sum (y - mean y)^2
    per_tag = {
        "explained_variance_score": lambda: explained(np.maximum(see / n - (se / n) ** 2, 0.0), tss / n),
        "r2_score": lambda: explained(see, tss),
        "mean_squared_error": lambda: see / n * s * s,
        "mean_absolute_error": lambda: sae / n * np.abs(s),
    }
    out = {}
    for name in metrics:
        if name not in per_tag:
            raise ValueError(f"metric {name!r} is not one of {MOMENT_METRICS}")
        values = per_tag[name]()
        out[name] = (values, values.mean(axis=-1))
    return out


def scores_block(moment_scores: Dict[str, Tuple[np.ndarray, np.ndarray]], tags: Sequence) -> dict:
    """``scores`` of the build metadata from ``scores_from_moments``: same keys and summaries as ModelBuilder writes."""
    out = {}
    for name, (per_tag, averaged) in moment_scores.items():
        label = name.replace("_", "-")
        # all tags at once (a thousand machines x 65 keys x 4 metrics is too many tiny numpy calls): rows = folds, last column = the average
        v = np.concatenate([np.asarray(per_tag, dtype=np.float64), np.asarray(averaged, dtype=np.float64)[:, None]], axis=1)
        stats = np.stack([v.mean(axis=0), v.std(axis=0), v.max(axis=0), v.min(axis=0)]).T.tolist()
        folds = v.T.tolist()
        keys = [f"{label}-{_column_label(tag)}" for tag in tags] + [label]
        for key, (mean, std, vmax, vmin), values in zip(keys, stats, folds):
            summary = {"fold-mean": mean, "fold-std": std, "fold-max": vmax, "fold-min": vmin}
            summary.update({f"fold-{i + 1}": x for i, x in enumerate(values)})
            out[key] = summary
    return out


def _get_data(dataset):
    if hasattr(dataset, "get_data"):
        X, y = dataset.get_data()
        meta = dataset.get_metadata() if hasattr(dataset, "get_metadata") else {}
    elif isinstance(dataset, dict) and "X" in dataset:
        X, y, meta = dataset["X"], dataset.get("y"), dataset.get("metadata", {})
    elif isinstance(dataset, (tuple, list)) and len(dataset) == 2:
        (X, y), meta = dataset, {}
    else:
        raise TypeError("dataset must provide get_data(), or be an (X, y) pair or {'X': ..., 'y': ...}; gordo's data providers "
                        "are outside this package")
    if not isinstance(X, pd.DataFrame):
        X = pd.DataFrame(np.asarray(X))
    if y is None:
        y = X
    if not isinstance(y, pd.DataFrame):
        y = pd.DataFrame(np.asarray(y), index=X.index)
    return X, y, meta


def _machine_dict(machine) -> dict:
    machine = machine.to_dict() if hasattr(machine, "to_dict") and not isinstance(machine, dict) else machine
    if "name" not in machine or "model" not in machine or "dataset" not in machine:
        raise ValueError("a machine needs at least 'name', 'model' and 'dataset'")
    return machine


def _machine_out(machine: dict, build_metadata: dict) -> dict:
    """The machine as ``Machine.to_dict()`` would give it after a build: the input plus ``metadata.build_metadata``."""
    out = {k: v for k, v in machine.items() if k != "dataset"}
    ds = machine["dataset"]
    out["dataset"] = ds.to_dict() if hasattr(ds, "to_dict") else (ds if isinstance(ds, dict) and "X" not in ds else {})
    meta = copy.deepcopy(machine.get("metadata") or {})
    meta.setdefault("user_defined", {})
    meta["build_metadata"] = build_metadata
    out["metadata"] = meta
    out["evaluation"] = {**DEFAULT_EVALUATION, **(machine.get("evaluation") or {})}
    return out


def _now() -> str:
    return str(datetime.datetime.now(datetime.timezone.utc).astimezone())


class ModelBuilder:
    """
    Build one machine: ``ModelBuilder(machine).build(output_dir)`` -> ``(model, machine_dict)``; the machine dict carries
    ``metadata.build_metadata.{model,dataset}`` exactly where the reference puts it (build_model.py:291-339).
    """

    def __init__(self, machine):
        self.machine = _machine_dict(machine)

    @property
    def gordo_version(self) -> str:
        return __version__

    @staticmethod
    def set_seed(seed: int):
        # the fit loops draw their shuffling seed and initial weights from numpy's global state
        np.random.seed(seed)
        random.seed(seed)

    def build(self, output_dir: Optional[str] = None):
        model, machine = self._build()
        if output_dir is not None:
            serializer.dump(model, output_dir, metadata=machine)
        return model, machine

    def _build(self):
        machine = self.machine
        evaluation = {**DEFAULT_EVALUATION, **(machine.get("evaluation") or {})}
        self.set_seed(int(evaluation.get("seed", 0)))

        t0 = time.time()
        X, y, dataset_meta = _get_data(machine["dataset"])
        query_sec = time.time() - t0
        model = serializer.from_definition(machine["model"])

        cv_sec, scores, splits = None, {}, {}
        cv_mode = str(evaluation["cv_mode"]).lower()
        if cv_mode in ("cross_val_only", "full_build") and hasattr(model, "predict"):
            t0 = time.time()
            scorers = build_metrics_dict(metrics_from_list(evaluation.get("metrics")), y, scaler=evaluation.get("scoring_scaler"))
            split_obj = serializer.from_definition(evaluation.get("cv", DEFAULT_CV))
            splits = build_split_dict(X, split_obj)
            kw = dict(X=X, y=y, scoring=scorers, return_estimator=True, cv=split_obj)
            cv = model.cross_validate(**kw) if hasattr(model, "cross_validate") else cross_validate(model, **kw)
            scores = {name: fold_summary(cv[f"test_{name}"]) for name in scorers}
            cv_sec = time.time() - t0
        cross_validation = {"scores": scores, "cv_duration_sec": cv_sec, "splits": splits}
        dataset_block = {"query_duration_sec": query_sec, "dataset_meta": dataset_meta}
        if cv_mode == "cross_val_only":
            return model, _machine_out(machine, {"model": {"cross_validation": cross_validation}, "dataset": dataset_block})

        t0 = time.time()
        model.fit(X, y)
        fit_sec = time.time() - t0
        model_block = {
            "model_offset": determine_offset(model, X),
            "model_creation_date": _now(),
            "model_builder_version": self.gordo_version,
            "model_training_duration_sec": fit_sec,
            "cross_validation": cross_validation,
            "model_meta": extract_metadata_from_model(model),
        }
        return model, _machine_out(machine, {"model": model_block, "dataset": dataset_block})


# ------------------------------------------------------------------------------------------------ the whole project at once
class _Canonical:
    """What ``FleetModelBuilder`` needs to know about a machine that can take the batched path."""

    def __init__(self, index, machine, model, spec, X, y, dataset_meta, query_sec, fit, n_splits, evaluation, input_scaler):
        self.index, self.machine, self.model, self.spec, self.input_scaler = index, machine, model, spec, input_scaler
        self.X, self.y, self.dataset_meta, self.query_sec = X, y, dataset_meta, query_sec
        self.fit, self.n_splits, self.evaluation = fit, n_splits, evaluation

    def bucket(self):
        s = self.spec
        return (tuple(s.dims), tuple(s.acts), tuple(float(v) for v in s.l1), tuple(sorted(s.adam.items())), tuple(s.metrics),
                len(self.X), self.fit["epochs"], self.fit["batch_size"], self.fit["shuffle"], self.n_splits, int(self.evaluation.get("seed", 0)),
                self.input_scaler)


def _default_minmax(scaler) -> bool:
    return type(scaler) is MinMaxScaler and tuple(scaler.feature_range) == (0, 1) and not getattr(scaler, "clip", False)


def _canonical(index, machine) -> Optional[_Canonical]:
    """The machine as a candidate for the batched path, or ``None`` with the reason logged."""
    from .machine.model.anomaly.diff import DiffBasedAnomalyDetector
    from .machine.model.factories.specs import FFNetSpec
    from .machine.model.models import KerasAutoEncoder

    def no(reason):
        logger.info("machine %s takes the per-machine path: %s", machine["name"], reason)
        return None

    evaluation = {**DEFAULT_EVALUATION, **(machine.get("evaluation") or {})}
    if str(evaluation["cv_mode"]).lower() != "full_build":
        return no(f"cv_mode {evaluation['cv_mode']}")
    if any(m.rpartition(".")[2] not in MOMENT_METRICS or ("." in m and not m.startswith("sklearn.metrics.")) for m in evaluation["metrics"]):
        return no("evaluation metrics beyond the four moment metrics")
    scoring = evaluation.get("scoring_scaler")
    if scoring:
        scoring = serializer.from_definition(scoring) if isinstance(scoring, (str, dict)) else scoring
        if not _default_minmax(scoring):
            return no("scoring_scaler is not a default MinMaxScaler")
    split_obj = serializer.from_definition(evaluation.get("cv", DEFAULT_CV))
    if type(split_obj) is not TimeSeriesSplit or split_obj.max_train_size is not None or split_obj.test_size is not None or split_obj.gap:
        return no("cv is not a plain TimeSeriesSplit")

    model = serializer.from_definition(machine["model"])
    if type(model) is not DiffBasedAnomalyDetector or model.window is not None or model.shuffle:
        return no("model is not a plain DiffBasedAnomalyDetector")
    if not _default_minmax(model.scaler):
        return no("detector scaler is not a default MinMaxScaler")
    ae, input_scaler = model.base_estimator, False
    if type(ae) is Pipeline and len(ae.steps) == 2 and _default_minmax(ae.steps[0][1]):
        ae, input_scaler = ae.steps[1][1], True  # Pipeline([MinMaxScaler(), KerasAutoEncoder]): gordo's example config
    if type(ae) is not KerasAutoEncoder:
        return no("base_estimator is not a KerasAutoEncoder, bare or behind one default MinMaxScaler")
    fit_args = ae.extract_supported_fit_args(ae.kwargs)
    if fit_args.get("validation_split") or fit_args.get("callbacks"):
        return no("validation_split / callbacks need the per-epoch loop")

    t0 = time.time()
    X, y, dataset_meta = _get_data(machine["dataset"])
    query_sec = time.time() - t0
    ae.kwargs.update({"n_features": X.shape[1], "n_features_out": y.shape[1]})
    spec = ae._build_spec()
    if not isinstance(spec, FFNetSpec):
        return no("not a feed-forward network")
    if len(X) != len(y) or len(X) // (split_obj.n_splits + 1) == 0:
        return no("too few rows for the CV folds")
    fit = {"epochs": int(fit_args.get("epochs", 1)), "batch_size": int(fit_args.get("batch_size") or 32), "shuffle": bool(fit_args.get("shuffle", True))}
    return _Canonical(index, machine, model, spec, X, y, dataset_meta, query_sec, fit, split_obj.n_splits, evaluation, input_scaler)


class FleetModelBuilder:
    """
    Build every machine of a project: ``FleetModelBuilder(machines).build(output_dir)`` -> ``[(model, machine_dict), ...]`` in
    input order, each written to ``<output_dir>/<name>/`` when ``output_dir`` is given.  Results per machine are what
    ``ModelBuilder`` gives (same detector attributes and metadata keys); only the launch count differs.
    """

    def __init__(self, machines: Sequence):
        self.machines = [_machine_dict(m) for m in machines]
        names = [m["name"] for m in self.machines]
        if len(set(names)) != len(names):
            raise ValueError("machine names must be unique")

    def shard(self, rank: int, world: int) -> "FleetModelBuilder":
        """
        This rank's contiguous block of the project (``fleet.partition``): machines are independent, so a multi-GPU build is
        ``FleetModelBuilder(machines).shard(rank, world).build(output_dir)`` in every process, with no collective at all.
        """
        from . import fleet

        return FleetModelBuilder([self.machines[i] for i in fleet.partition(len(self.machines), world)[rank]])

    def build(self, output_dir: Optional[str] = None) -> List[Tuple[Any, dict]]:
        results: List[Optional[Tuple[Any, dict]]] = [None] * len(self.machines)
        buckets: Dict[tuple, List[_Canonical]] = {}
        for i, machine in enumerate(self.machines):
            c = _canonical(i, machine)
            if c is None:
                results[i] = ModelBuilder(machine).build()
            else:
                buckets.setdefault(c.bucket(), []).append(c)
        for members in buckets.values():
            try:
                built_bucket = self._build_bucket(members)
            except Exception as exc:  # e.g. an architecture the batched fit kernel cannot hold: the machines still get built, one by one
                logger.warning("batched build of %d machines failed (%s: %s); building them one at a time", len(members), type(exc).__name__, exc)
                built_bucket = [ModelBuilder(c.machine).build() for c in members]
            for c, built in zip(members, built_bucket):
                results[c.index] = built
        if output_dir is not None:
            for model, machine in results:
                serializer.dump(model, os.path.join(output_dir, machine["name"]), metadata=machine)
        return results

    @staticmethod
    def _build_bucket(members: List[_Canonical]) -> List[Tuple[Any, dict]]:
        from . import engine, fleet

        first = members[0]
        eng = engine.ff_engine_for(first.spec)
        rows, K = len(first.X), first.n_splits
        t0 = time.time()
        x_host = np.concatenate([np.ascontiguousarray(c.X.values, dtype=np.float32) for c in members])
        same_y = all(c.y is c.X for c in members)
        xd = engine.to_device_f32(x_host, eng.device)
        yd = xd if same_y else engine.to_device_f32(np.concatenate([np.ascontiguousarray(c.y.values, dtype=np.float32) for c in members]), eng.device)
        fb = fleet.build_fleet(eng, xd, yd, rows, epochs=first.fit["epochs"], batch_size=first.fit["batch_size"], n_splits=K,
                               seed=int(first.evaluation.get("seed", 0)), adam=first.spec.adam, shuffle=first.fit["shuffle"],
                               input_scaler=first.input_scaler)
        moments = fb.cv_moments.cpu().numpy()
        scale = fb.scale.cpu().numpy().astype(np.float64)
        engine._torch().cuda.synchronize()
        share = (time.time() - t0) / len(members)  # the bucket's wall time, spread evenly: there is no per-machine time any more
        test = rows // (K + 1)
        split_obj = TimeSeriesSplit(n_splits=K)
        out = []
        for m, c in enumerate(members):
            tags = list(c.y.columns)
            model = fb.detector(m, tags=tags, template=c.model, input_tags=list(c.X.columns))
            names = [s.rpartition(".")[2] for s in c.evaluation["metrics"]]
            scoring_scale = scale[m] if c.evaluation.get("scoring_scaler") else None
            scores = scores_block(scores_from_moments(moments[m], test, scoring_scale, names), tags)
            model_block = {
                "model_offset": 0,  # a Dense stack answers every row
                "model_creation_date": _now(),
                "model_builder_version": __version__,
                "model_training_duration_sec": share * 1.0 / (K + 1),
                "cross_validation": {"scores": scores, "cv_duration_sec": share * K / (K + 1), "splits": build_split_dict(c.X, split_obj)},
                "model_meta": extract_metadata_from_model(model),
            }
            dataset_block = {"query_duration_sec": c.query_sec, "dataset_meta": c.dataset_meta}
            out.append((model, _machine_out(c.machine, {"model": model_block, "dataset": dataset_block})))
        return out


# ------------------------------------------------------------------------------------------------ from a project config
MACHINE_YAML_FIELDS = ("model", "dataset", "evaluation", "metadata", "runtime")  # may arrive as YAML text blocks (machine/constants.py)


def patch_dict(original: dict, patch: dict) -> dict:
    """
    ``original`` with every path of ``patch`` added or replaced, nothing removed (workflow_generator/helpers.py:16-45).  Lists are
    replaced as a whole; the reference patches through dictdiffer [3P, not installed here], which walks lists element by element.
    """
    out = copy.deepcopy(original)
    for key, value in (patch or {}).items():
        if isinstance(value, dict) and isinstance(out.get(key), dict):
            out[key] = patch_dict(out[key], value)
        else:
            out[key] = copy.deepcopy(value)
    return out


class RandomDataset:
    """
    Stand-in for gordo-core's ``RandomDataProvider`` datasets [3P, not installed]: seeded uniform noise for the configured tags on
    the regular ``resolution`` grid between ``train_start_date`` and ``train_end_date``.  It exists so that project configs written
    for the reference's tests and docs (``data_provider: {type: RandomDataProvider}``) build here; real data comes from any object
    with ``get_data()`` passed through ``datasets=``.
    """

    def __init__(self, **config):
        self.config = config
        tags = config.get("tag_list") or config.get("tags")
        if not tags:
            raise ValueError("dataset needs 'tag_list' (or 'tags')")
        self.tag_list = [t["name"] if isinstance(t, dict) else str(t) for t in tags]
        targets = config.get("target_tag_list") or tags
        self.target_tag_list = [t["name"] if isinstance(t, dict) else str(t) for t in targets]
        self.resolution = config.get("resolution", "10min")
        self.start, self.end = pd.Timestamp(config["train_start_date"]), pd.Timestamp(config["train_end_date"])
        if self.start.tzinfo is None or self.end.tzinfo is None:
            raise ValueError("train_start_date / train_end_date need a timezone")
        if self.start >= self.end:
            raise ValueError(f"train_end_date ({self.end}) must be after train_start_date ({self.start})")

    def get_data(self):
        import zlib

        index = pd.date_range(self.start, self.end, freq=pd.tseries.frequencies.to_offset(self.resolution), inclusive="left")
        names = list(dict.fromkeys(self.tag_list + self.target_tag_list))
        rng = np.random.default_rng(zlib.crc32("|".join(names).encode()))
        data = pd.DataFrame(rng.random((len(index), len(names))), index=index, columns=names)
        return data[self.tag_list], data[self.target_tag_list]

    def get_metadata(self):
        return {"tag_list": self.tag_list, "target_tag_list": self.target_tag_list, "resolution": self.resolution,
                "train_start_date": str(self.start), "train_end_date": str(self.end)}

    def to_dict(self):
        out = {k: v for k, v in self.config.items() if k != "tags"}
        out.update({"type": "RandomDataset", "tag_list": self.tag_list, "target_tag_list": self.target_tag_list, "resolution": self.resolution})
        return out


def _dataset_from_config(config: dict):
    provider = (config.get("data_provider") or {}).get("type", "") if isinstance(config.get("data_provider"), dict) else ""
    if str(config.get("type", "")).endswith("RandomDataset") or str(provider).endswith("RandomDataProvider"):
        return RandomDataset(**config)
    raise TypeError("only RandomDataset / RandomDataProvider dataset configs can be instantiated here; pass datasets= (name or machine -> an "
                    "object with get_data()) for real data")


def machines_from_config(config, project_name: str = "local-build", datasets=None) -> List[dict]:
    """
    The machines of a project config -- ``{"machines": [...], "globals": {...}}`` as a dict or YAML text -- in the dict form the
    builders take, with the globals folded in the way ``Machine.from_config`` does (gordo/machine/machine.py:78-149: the machine's
    model wins, runtime and evaluation are globals patched by the machine, dataset is the machine patched by the globals) over
    the default evaluation of ``NormalizedConfig``.  ``datasets``: mapping name -> dataset object, or a callable taking the machine.
    """
    import yaml

    if isinstance(config, str):
        config = yaml.safe_load(config)
    if not isinstance(config, dict) or not config.get("machines"):
        raise ValueError("config needs a non-empty 'machines' list")

    def parsed(block: dict) -> dict:
        out = dict(block or {})
        for field in MACHINE_YAML_FIELDS:
            if isinstance(out.get(field), str):
                out[field] = yaml.safe_load(out[field])
        return out

    config_globals = patch_dict({"evaluation": DEFAULT_EVALUATION}, parsed(config.get("globals")))
    machines = []
    for conf in config["machines"]:
        conf = parsed(conf)
        if "name" not in conf:
            raise ValueError("every machine needs a name")
        model = conf.get("model") or config_globals.get("model")
        if model is None:
            raise ValueError("model is empty")
        machine = {
            "name": conf["name"],
            "project_name": conf.get("project_name") or project_name,
            "model": model,
            "runtime": patch_dict(config_globals.get("runtime") or {}, conf.get("runtime") or {}),
            "evaluation": patch_dict(config_globals.get("evaluation") or {}, conf.get("evaluation") or {"cv_mode": "full_build"}),
            "metadata": {"user_defined": {"global-metadata": config_globals.get("metadata") or {}, "machine-metadata": conf.get("metadata") or {}}},
        }
        dataset_config = patch_dict(conf.get("dataset") or {}, config_globals.get("dataset") or {})
        if callable(datasets):
            machine["dataset"] = datasets({**machine, "dataset": dataset_config})
        elif datasets is not None and conf["name"] in datasets:
            machine["dataset"] = datasets[conf["name"]]
        else:
            machine["dataset"] = _dataset_from_config(dataset_config)
        machines.append(machine)
    return machines


def local_build(config_str, datasets=None, batched: bool = True):
    """
    Build the model(s) of a bare gordo config locally and yield ``(model, machine)`` per machine, in config order
    (gordo/builder/local_build.py:15-80).  ``batched=False`` builds one machine at a time like the reference does.
    """
    machines = machines_from_config(config_str, datasets=datasets)
    if batched:
        yield from FleetModelBuilder(machines).build()
    else:
        for machine in machines:
            yield ModelBuilder(machine).build()
