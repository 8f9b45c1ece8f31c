"""Protocol every gordo model exposes (mirror of gordo/machine/model/base.py:10-35)."""
import abc


class GordoBase(abc.ABC):
    @abc.abstractmethod
    def __init__(self, **kwargs):
        ...

    @abc.abstractmethod
    def get_params(self, deep=False):
        """Parameters the object was constructed with."""

    @abc.abstractmethod
    def score(self, X, y, sample_weight=None):
        """Default scorer of the model type."""

    @abc.abstractmethod
    def get_metadata(self):
        """Model specific metadata, if any."""
