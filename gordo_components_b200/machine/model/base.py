"""Protocol every gordo model exposes (mirror of gordo/machine/model/base.py:10-35)."""
import abc
import importlib


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


def register_with_gordo(module: str, name: str, cls) -> bool:
    """
    gordo's own callers test ``isinstance(obj, gordo.machine.model.base.GordoBase)`` -- ``ModelBuilder._extract_metadata_from_model``
    (gordo/builder/build_model.py:552-553, 566) only collects ``get_metadata()`` of objects that pass it.  Where gordo is installed,
    the protocol classes here are therefore registered as virtual subclasses of gordo's ABCs, so the estimators of this package are
    ``GordoBase`` instances to gordo's builder and serializer without inheriting from (or importing anything else of) gordo.
    Returns whether the registration happened; without gordo there is nothing to register with.
    """
    try:
        ref = getattr(importlib.import_module(module), name)
    except Exception:  # gordo absent, or present without its own dependencies
        return False
    if ref is not cls and isinstance(ref, abc.ABCMeta):
        ref.register(cls)
        return True
    return False


register_with_gordo("gordo.machine.model.base", "GordoBase", GordoBase)
