"""CPU checks of the boundary: libpkrylov.so loads, exports every symbol include/pkrylov.h declares, the ctypes
table covers them all, the product path fails loudly without a GPU and never imports the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pkrylov.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pk_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from parallel_krylov_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pkrylov.h but not exported by libpkrylov.so"
    assert set(_lib.exported_symbols()) == set(names), set(_lib.exported_symbols()) ^ set(names)
    assert lib.pk_version() == 102
    assert lib.pk_work_doubles(0, 1024, 0) == 3 * 1024          # cg: r, p, v
    assert lib.pk_work_doubles(3, 1024, 8) == (10 + 9 + 3) * 1024   # Ar, Ay, z, spare Ar0, A r


def test_struct_layouts_match_header():
    from parallel_krylov_b200._lib import SolveOpts, SolveResult
    assert ctypes.sizeof(SolveOpts) == 72 and ctypes.sizeof(SolveResult) == 56


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_fails_loudly_without_gpu():
    import parallel_krylov_b200 as pk
    with pytest.raises(pk.PkError):
        pk.cg(np.eye(4), np.ones(4))
    with pytest.raises(pk.PkError):
        pk.Operator.from_any(np.eye(4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "parallel_krylov_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "krylov_oracle" not in text and "host_kernels" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
