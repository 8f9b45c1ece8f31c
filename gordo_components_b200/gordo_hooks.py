"""
The hooks gordo itself offers for swapping its builder (SURVEY section 8 b / f1): ``gordo build --model-builder-class`` /
``MODEL_BUILDER_CLASS`` (gordo/cli/cli.py:81-86, 147-148) imports a class path and requires a subclass of
``gordo.builder.build_model.ModelBuilder`` (gordo/builder/utils.py:8-17).

    MODEL_BUILDER_CLASS=gordo_components_b200.gordo_hooks.B200ModelBuilder gordo build ...

``B200ModelBuilder`` IS gordo's ModelBuilder -- same ``build`` / ``_build`` / metadata / cache-key code, inherited -- with the one
method that assumes TensorFlow replaced: ``set_seed`` seeds NumPy and ``random`` (which is where this package's initial weights and
shuffling keys come from) and TensorFlow only if it is importable.  The class can only exist where gordo is installed, so it is
created on first access; without gordo the attribute lookup raises ImportError with that explanation.  (The model classes themselves
need no builder hook: they drop into gordo's stock ModelBuilder through their class paths in the model definition.)
"""
from __future__ import annotations

import importlib
import random

import numpy as np

_cache = {}


def _make():
    try:
        base = importlib.import_module("gordo.builder.build_model").ModelBuilder
    except Exception as e:  # gordo (or one of its own dependencies) is not importable here
        raise ImportError(f"B200ModelBuilder subclasses gordo.builder.build_model.ModelBuilder, which cannot be imported: {e}") from e

    class B200ModelBuilder(base):
        """gordo's ModelBuilder; seeds NumPy / random (this package's initial weights and shuffle keys), TensorFlow only if present."""

        def set_seed(self, seed: int):
            try:
                importlib.import_module("tensorflow").random.set_seed(seed)
            except Exception:
                pass
            np.random.seed(seed)
            random.seed(seed)

    B200ModelBuilder.__module__ = __name__
    return B200ModelBuilder


def __getattr__(name):
    if name == "B200ModelBuilder":
        if name not in _cache:
            _cache[name] = _make()
        return _cache[name]
    raise AttributeError(name)
