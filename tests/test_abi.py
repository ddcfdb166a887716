"""CPU: the C-ABI library loads and exports every symbol include/hiarch_b200.h declares
(no compute calls without a GPU)."""
import os
import re

from hiarch_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hiarch_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hia_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from hiarch_b200 import build
    build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hiarch_b200.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype in hiarch_b200/_lib.py"
    for name in _lib.PROTOTYPES:
        assert name in declared, f"{name} bound in _lib.py but not declared in the header"
    assert lib.hia_version() >= 100
    assert lib.hia_status_string(0) == b"ok"
    assert lib.hia_status_string(-5) == b"workspace too small"


def test_ops_fail_loudly_without_cuda_tensors():
    import pytest
    import torch
    from hiarch_b200 import ops
    with pytest.raises(_lib.HiaError):
        ops.similarity(torch.zeros(8, 8))
