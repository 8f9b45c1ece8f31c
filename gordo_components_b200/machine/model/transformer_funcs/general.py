"""Functions for ``sklearn.preprocessing.FunctionTransformer`` steps of a model definition (gordo/machine/model/transformer_funcs/general.py:22-26)."""


def multiply_by(X, factor):
    """``X * factor``: the per-tag affine step production configs put in front of the auto-encoder."""
    return X * factor
