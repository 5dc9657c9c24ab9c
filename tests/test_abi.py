"""CPU: the C-ABI library builds/loads and exports every symbol include/sgcn_b200.h declares, and
argument validation works without touching a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "sgcn_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(sgcn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from stochastic_gcn_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libsgcn_b200.so does not export %s" % n
    assert set(names) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with the header"
    assert lib.sgcn_abi_version() == 1


def test_argument_validation_needs_no_gpu():
    from stochastic_gcn_b200 import _lib
    lib = _lib.load()
    # negative sizes / null pointers are rejected before any CUDA call
    rc = lib.sgcn_gather_rows(None, 4, None, -1, None, 4, None, 4, None)
    assert rc == _lib.SGCN_EINVAL and b"negative" in lib.sgcn_last_error()
    rc = lib.sgcn_gather_rows(None, 4, None, 3, None, 4, None, 4, None)
    assert rc == _lib.SGCN_EINVAL
    rc = lib.sgcn_spmm_csr(None, None, None, None, 5, None, None, 2, 8, None, 8, 0, None)
    assert rc == _lib.SGCN_EINVAL
    rc = lib.sgcn_sampler_expand(None, 1, 0)
    assert rc == _lib.SGCN_EINVAL
    # zero-sized work is a no-op success
    assert lib.sgcn_gather_rows(None, 4, None, 0, None, 4, None, 4, None) == 0
    assert lib.sgcn_spmm_coo(None, None, 0, None, 4, 4, None, 4, 0, None) == 0
    with pytest.raises(_lib.SgcnError):
        _lib.check(lib.sgcn_history_update(None, 1, None, 2, None, None, 1, 4, None, None))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stochastic_gcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "sgcn_oracle" not in text and "libsgcn_ref" not in text, f
