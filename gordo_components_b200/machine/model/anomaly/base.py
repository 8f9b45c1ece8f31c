"""Anomaly detector protocol (mirror of gordo/machine/model/anomaly/base.py:11-23)."""
import abc
from datetime import timedelta
from typing import Optional

import pandas as pd
from sklearn.base import BaseEstimator

from ..base import GordoBase, register_with_gordo


class AnomalyDetectorBase(BaseEstimator, GordoBase, metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def anomaly(self, X: pd.DataFrame, y: pd.DataFrame, frequency: Optional[timedelta] = None) -> pd.DataFrame:
        """Frame of model output and anomaly scores for ``X`` against ``y``."""


register_with_gordo("gordo.machine.model.anomaly.base", "AnomalyDetectorBase", AnomalyDetectorBase)
