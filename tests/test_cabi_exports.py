"""The C-ABI library loads on a GPU-less box and exports every symbol include/gordo_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from gordo_components_b200 import _cabi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    return _cabi.load_library()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "gordo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gordo_b200.h but not exported"
    assert set(names) == set(_cabi.EXPORTS), "ctypes binding and header disagree on the entry points"


def test_abi_version_and_struct_layout(lib):
    assert lib.gb_abi_version() == 2
    assert C.sizeof(_cabi.GbJob) == 24
    assert C.sizeof(_cabi.GbFFNet) == 4 + 4 * 17 + 4 * 16 + 4 * 16
    assert C.sizeof(_cabi.GbFitHParams) == 48
    assert C.sizeof(_cabi.GbLstmNet) == 4 * 3 + 4 * 16 * 2 + 8


def test_param_counts_match_survey(lib):
    hg = lambda t, dims: _cabi.make_ffnet([t, *dims, *dims[::-1], t], ["tanh"] * 6 + ["linear"])  # noqa: E731
    assert lib.gb_ffnet_param_count(C.byref(hg(64, (53, 43, 32)))) == 15438
    assert lib.gb_ffnet_param_count(C.byref(hg(8, (7, 5, 4)))) == 278
    assert lib.gb_ffnet_param_stride(C.byref(hg(8, (7, 5, 4)))) == 280
    assert lib.gb_ffnet_param_count(C.byref(hg(128, (107, 85, 64)))) == 61198
    ls = _cabi.make_lstmnet(128, [256, 128, 64, 64, 128, 256], ["tanh"] * 6, 128, "linear", 144)
    assert lib.gb_lstm_param_count(C.byref(ls)) == 1199744


def test_argument_validation_needs_no_gpu(lib):
    net = _cabi.make_ffnet([4, 3, 4], ["tanh", "linear"])
    rc = lib.gb_ffae_infer_score(C.byref(net), None, None, 1, 10, 10, 10, None, None, None, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"non-NULL" in lib.gb_last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc)
    bad = _cabi.make_ffnet([4, 3, 4], ["tanh", "linear"])
    bad.dims[1] = 4096
    assert lib.gb_ffnet_param_count(C.byref(bad)) == 0
    with pytest.raises(ValueError):
        _cabi.make_ffnet([4, 3, 4], ["swish", "linear"])


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np

    from gordo_components_b200.machine.model.models import KerasAutoEncoder

    with pytest.raises(_cabi.GordoB200Error):
        KerasAutoEncoder(kind="feedforward_hourglass").fit(np.random.rand(16, 4), np.random.rand(16, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gordo_components_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
