"""
``InfImputer``: replaces +/-inf in sensor data before it reaches the network (gordo/machine/model/transformers/imputer.py:12-127).
Host-side preprocessing -- one masked pass per call, nothing for the GPU to win here.
"""
import numpy as np
import pandas as pd
from sklearn.base import TransformerMixin


class InfImputer(TransformerMixin):
    def __init__(self, inf_fill_value=None, neg_inf_fill_value=None, strategy="minmax", delta: float = 2.0):
        """
        strategy ``"minmax"``: +inf -> (largest finite value of that feature seen in ``fit``) + delta, -inf -> smallest - delta,
        clipped to the dtype's range; ``"extremes"``: the dtype's max / min; ``None``: only the explicit fill values apply.
        Explicit ``inf_fill_value`` / ``neg_inf_fill_value`` win over the strategy.
        """
        self.inf_fill_value = inf_fill_value
        self.neg_inf_fill_value = neg_inf_fill_value
        self.strategy = strategy
        self.delta = delta

    def get_params(self, deep=True):
        return {"inf_fill_value": self.inf_fill_value, "neg_inf_fill_value": self.neg_inf_fill_value, "strategy": self.strategy,
                "delta": self.delta}

    def set_params(self, **params):
        for k, v in params.items():
            setattr(self, k, v)
        return self

    def fit(self, X, y=None):
        if self.strategy == "minmax":
            values = np.asarray(pd.DataFrame(X).values)
            if not np.issubdtype(values.dtype, np.floating):
                values = values.astype(np.float64)
            info = np.finfo(values.dtype)
            finite = np.isfinite(values)
            with np.errstate(invalid="ignore"):
                hi = np.where(finite, values, -np.inf).max(axis=0)
                lo = np.where(finite, values, np.inf).min(axis=0)
            hi = np.where(finite.any(axis=0), hi, np.nan)  # a feature with no finite sample has nothing to anchor on
            lo = np.where(finite.any(axis=0), lo, np.nan)
            self._posinf_fill_values = np.where(info.max - self.delta > hi, hi + self.delta, info.max)
            self._neginf_fill_values = np.where(info.min + self.delta < lo, lo - self.delta, info.min)
        return self

    def transform(self, X, y=None):
        X = X.values if isinstance(X, pd.DataFrame) else X
        if not X.flags.writeable:  # pandas copy-on-write hands out read-only views
            X = X.copy()
        if self.inf_fill_value is not None:
            X[np.isposinf(X)] = self.inf_fill_value
        if self.neg_inf_fill_value is not None:
            X[np.isneginf(X)] = self.neg_inf_fill_value
        if self.strategy is None:
            return X
        if self.strategy == "extremes":
            info = np.finfo(X.dtype)
            X[np.isposinf(X)] = info.max
            X[np.isneginf(X)] = info.min
            return X
        if self.strategy == "minmax":
            pos, neg = np.isposinf(X), np.isneginf(X)
            if pos.any():
                X[pos] = np.broadcast_to(self._posinf_fill_values, X.shape)[pos]
            if neg.any():
                X[neg] = np.broadcast_to(self._neginf_fill_values, X.shape)[neg]
            return X
        raise AttributeError(f"unknown InfImputer strategy {self.strategy!r}")
