"""The C-ABI shared library loads without a GPU and exports every symbol include/resuneta.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dll():
    import __graft_entry__ as g
    g.build()
    from resuneta_b200 import _capi
    return _capi.load_cdll()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "resuneta.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rsa_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_are_exported(dll):
    names = _declared()
    assert len(names) >= 29
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/resuneta.h but not exported"


def test_binding_table_matches_header(dll):
    from resuneta_b200 import _capi
    assert sorted(_capi.EXPORTS) == _declared()


def test_version_and_error_string(dll):
    assert b"sm_100a" in dll.rsa_version()
    assert isinstance(dll.rsa_last_error(), bytes)


def test_argument_validation_needs_no_device(dll):
    # shape errors are reported before any CUDA call
    rc = dll.rsa_softmax_fwd(None, None, 0, 0, None)
    assert rc == -1 and b"softmax_fwd" in dll.rsa_last_error()
    rc = dll.rsa_adam_step(None, None, None, None, 0, 0.0, None, 0.9, 0.999, 1e-7, 1.0, None)
    assert rc == -1


def test_no_cpu_fallback_without_device():
    import torch
    from resuneta_b200 import _capi
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(RuntimeError):
        _capi.Lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "resunet-a_mltsk_keras_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle\.", src, re.M), f"{f} uses the oracle"
