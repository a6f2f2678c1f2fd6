"""CPU: the C-ABI library loads without a GPU, exports every symbol include/msig.h declares, the
ctypes table covers them, and it fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

import msig_b200  # noqa: F401
from msig_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "msig.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(msig_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    dll = ctypes.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, f"declared in msig.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    declared = set(_declared())
    table = set(lib.exported_symbols())
    assert table <= declared, f"bound but not declared: {sorted(table - declared)}"
    assert declared <= table, f"declared but not bound: {sorted(declared - table)}"


def test_version_and_loud_failure_without_gpu():
    import torch
    l = lib.load()
    assert l.msig_version() >= 100
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU path"):
        lib.init(0)


def test_modules_refuse_cpu_tensors():
    import torch
    from msig_b200 import model
    G = model.StyleCycleGANGenerator()
    with pytest.raises(RuntimeError, match="no CPU path"):
        G(torch.zeros(1, 3, 64, 64), torch.zeros(1, 256))
    D = model.MultiDomainDiscriminator(num_domains=3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        D(torch.zeros(1, 3, 64, 64), None)
