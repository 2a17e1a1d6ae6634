"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/jstacs_hmm.h declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "jstacs_hmm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jh_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native

    return _native


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jstacs_hmm.h but not exported"
    assert built.lib().jh_abi_version() == built.ABI_VERSION


def test_binding_signatures_cover_the_header(built):
    built.lib()
    for n in declared_symbols():
        assert getattr(built.lib(), n).argtypes is not None, f"{n} is not bound in _native.py"


def test_no_cpu_fallback_without_a_device(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from jstacs_b200.workloads import ergodic_hmm, make_data

    hmm = ergodic_hmm(2, 1)
    with pytest.raises(built.NativeError) as ei:
        hmm.getLogScoreFor(make_data(2, 10))
    assert ei.value.status == -3 and "no CPU fallback" in str(ei.value)


def test_oracle_is_not_reachable_from_the_product():
    """The product package must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "jstacs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "hmm_oracle", "jo_model", "oracle/", "oracle."):
                    assert needle not in src, f"{f} references the oracle ({needle})"
    out = os.popen(f"ldd {os.path.join(pkg, 'lib', 'libjstacs_hmm_b200.so')} 2>/dev/null").read()
    assert "oracle" not in out
