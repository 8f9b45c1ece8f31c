"""
TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the parts of the *unmodified* reference that can execute in this container
straight from ``/root/reference`` (read-only), so that golden vectors under
``tests/golden/`` are produced by the reference's own code, not by our restatement:

* ``gordo/machine/model/anomaly/diff.py``  (DiffBasedAnomalyDetector, KFCV variant)
* ``gordo/machine/model/utils.py``         (make_base_dataframe, metric_wrapper)
* ``gordo/machine/model/factories/utils.py`` (hourglass_calc_dims, check_dim_func_len)

TensorFlow / Keras / scikeras / xarray / gordo_core are not installed here, so the
modules they would provide are replaced by inert stubs *before* import; none of the
stubbed symbols take part in the anomaly arithmetic.  ``/root/reference`` does not
exist on the GPU box: everything that runs there uses the committed fixtures instead.

Nothing is copied: the reference files are executed where they lie.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GORDO_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, "gordo", "machine", "model", "anomaly", "diff.py")
    )


def _ns_module(name: str, path: str | None = None) -> types.ModuleType:
    mod = types.ModuleType(name)
    if path is not None:
        mod.__path__ = [path]  # namespace-style package; its __init__.py is skipped
    sys.modules[name] = mod
    return mod


def _install_stubs() -> None:
    """Inert stand-ins for third-party packages the reference imports at module scope."""
    if "tensorflow" not in sys.modules:
        tf = _ns_module("tensorflow", path="<stub>")
        keras = _ns_module("tensorflow.keras", path="<stub>")
        tf.keras = keras
        for sub in ("models", "layers", "optimizers", "callbacks", "regularizers"):
            m = _ns_module(f"tensorflow.keras.{sub}")
            setattr(keras, sub, m)
        keras.models.Model = type("Model", (), {})
        keras.models.Sequential = type("Sequential", (), {})
        keras.models.load_model = lambda *a, **k: None
        keras.models.save_model = lambda *a, **k: None
        keras.optimizers.Optimizer = type("Optimizer", (), {})
        keras.layers.Dense = type("Dense", (), {})
        keras.layers.LSTM = type("LSTM", (), {})
        pre = _ns_module("tensorflow.keras.preprocessing", path="<stub>")
        seq = _ns_module("tensorflow.keras.preprocessing.sequence")
        seq.pad_sequences = lambda *a, **k: None
        seq.TimeseriesGenerator = type("TimeseriesGenerator", (), {})
        pre.sequence = seq
        keras.preprocessing = pre
    if "keras" not in sys.modules:
        k = _ns_module("keras", path="<stub>")
        ks = _ns_module("keras.src", path="<stub>")
        kc = _ns_module("keras.src.callbacks")
        kc.Callback = type("Callback", (), {})
        k.src = ks
        ks.callbacks = kc
    if "scikeras" not in sys.modules:
        sk = _ns_module("scikeras", path="<stub>")
        w = _ns_module("scikeras.wrappers")

        class KerasRegressor:  # noqa: D401 - stub
            _fit_kwargs: set = set()
            _predict_kwargs: set = set()
            _compile_kwargs: set = set()
            model = None

            def __init__(self, **kwargs):
                pass

            def get_params(self, **kw):
                return {}

        w.KerasRegressor = KerasRegressor
        sk.wrappers = w
    if "xarray" not in sys.modules:
        xr = _ns_module("xarray")
        xr.DataArray = type("DataArray", (), {})
        xr.Dataset = type("Dataset", (), {})
    if "simplejson" not in sys.modules:
        import json

        sj = _ns_module("simplejson")
        sj.dumps, sj.loads, sj.dump, sj.load = json.dumps, json.loads, json.dump, json.load
    if "gordo_core" not in sys.modules:
        gc = _ns_module("gordo_core", path="<stub>")
        st = _ns_module("gordo_core.sensor_tag")

        class SensorTag:  # minimal value object; only `.name` is read by model/utils.py
            def __init__(self, name, **kw):
                self.name = name

        st.SensorTag = SensorTag
        iu = _ns_module("gordo_core.import_utils")
        iu.import_location = lambda loc: importlib.import_module(loc)
        gc.sensor_tag, gc.import_utils = st, iu


def _install_pandas_append_shim() -> None:
    """
    The reference pins pandas 1.5.3 and calls ``DataFrame.append`` (diff.py:235-237,
    :250-253); pandas >= 2 removed it.  Same semantics, expressed with concat.
    """
    import pandas as pd

    if not hasattr(pd.DataFrame, "append"):

        def _append(self, other, ignore_index=False):
            if isinstance(other, pd.Series):
                other = other.to_frame().T
            if self.empty and len(self.columns) == 0:
                return other.copy()
            return pd.concat([self, other], ignore_index=ignore_index)

        pd.DataFrame.append = _append  # type: ignore[attr-defined]


_loaded: dict = {}


def load_reference():
    """
    Returns a namespace with the reference's own objects:
    ``DiffBasedAnomalyDetector, DiffBasedKFCVAnomalyDetector, make_base_dataframe,
    metric_wrapper, hourglass_calc_dims, check_dim_func_len``.
    """
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise FileNotFoundError(
            f"reference tree not found under {REFERENCE_ROOT}; use tests/golden fixtures"
        )
    _install_stubs()
    _install_pandas_append_shim()
    g = os.path.join(REFERENCE_ROOT, "gordo")
    # package skeleton whose __init__ chains (gordo_core, dataclasses_json ...) are skipped
    for name, sub in (
        ("gordo", ""),
        ("gordo.machine", "machine"),
        ("gordo.machine.model", "machine/model"),
        ("gordo.machine.model.anomaly", "machine/model/anomaly"),
        ("gordo.machine.model.factories", "machine/model/factories"),
    ):
        if name not in sys.modules:
            _ns_module(name, os.path.join(g, sub))
    # `from gordo import serializer` in models.py -- only attribute access at call time
    if "gordo.serializer" not in sys.modules:
        ser = _ns_module("gordo.serializer")
        sys.modules["gordo"].serializer = ser
    # the real models.py cannot import (it needs TF); diff.py only needs the class
    # object for its default argument, which our goldens never use.
    if "gordo.machine.model.models" not in sys.modules:
        mm = _ns_module("gordo.machine.model.models")

        class KerasAutoEncoder:  # placeholder for diff.py's default base_estimator
            def __init__(self, kind=None, **kw):
                self.kind = kind

        mm.KerasAutoEncoder = KerasAutoEncoder

    def _exec(modname: str, relpath: str):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(g, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    base = _exec("gordo.machine.model.base", "machine/model/base.py")
    mutils = _exec("gordo.machine.model.utils", "machine/model/utils.py")
    sys.modules["gordo.machine.model"].utils = mutils
    sys.modules["gordo.machine.model"].base = base
    _exec("gordo.machine.model.anomaly.base", "machine/model/anomaly/base.py")
    diff = _exec("gordo.machine.model.anomaly.diff", "machine/model/anomaly/diff.py")
    futils = _exec("gordo.machine.model.factories.utils", "machine/model/factories/utils.py")
    _loaded.update(
        DiffBasedAnomalyDetector=diff.DiffBasedAnomalyDetector,
        DiffBasedKFCVAnomalyDetector=diff.DiffBasedKFCVAnomalyDetector,
        make_base_dataframe=mutils.make_base_dataframe,
        metric_wrapper=mutils.metric_wrapper,
        hourglass_calc_dims=futils.hourglass_calc_dims,
        check_dim_func_len=futils.check_dim_func_len,
        GordoBase=base.GordoBase,
    )
    return types.SimpleNamespace(**_loaded)
