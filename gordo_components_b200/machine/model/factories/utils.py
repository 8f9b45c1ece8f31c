"""Layer-size helpers (behaviour of gordo/machine/model/factories/utils.py:7-63)."""
import math
from typing import Tuple


def hourglass_calc_dims(compression_factor: float, encoding_layers: int, n_features: int) -> Tuple[int, ...]:
    """
    Widths of the ``encoding_layers`` encoder layers of an hourglass network: a straight line
    from ``n_features`` down to ``ceil(compression_factor * n_features)`` (at least 1), each
    point rounded with Python's round-half-to-even.
    """
    if not (0 <= compression_factor <= 1):
        raise ValueError("compression_factor must be 0 <= compression_factor <= 1")
    if encoding_layers < 1:
        raise ValueError("encoding_layers must be >= 1")
    narrowest = max(min(math.ceil(compression_factor * n_features), n_features), 1)
    step = (n_features - narrowest) / encoding_layers
    return tuple(round(n_features - layer * step) for layer in range(1, encoding_layers + 1))


def check_dim_func_len(prefix: str, dim: Tuple[int, ...], func: Tuple[str, ...]):
    """One activation per layer, or ValueError."""
    if len(dim) != len(func):
        raise ValueError(
            f"The length (i.e. the number of network layers) of {prefix}_dim ({len(dim)}) and {prefix}_func "
            f"({len(func)}) must be equal. If only {prefix}_dim or {prefix}_func was passed, ensure that its "
            f"length matches that of the {prefix} parameter not passed."
        )
