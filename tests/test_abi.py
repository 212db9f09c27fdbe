"""The C-ABI library loads on a box without a GPU, exports every symbol include/wmd_b200.h
declares, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "wmd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wmd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from consistent__style_transfer_b200 import _lib
    L = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/wmd_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names
    assert b"sm_100a" in L.wmd_version()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from consistent__style_transfer_b200 import _lib
    from consistent__style_transfer_b200.engine import WMDEngine
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback|CUDA"):
        WMDEngine(np.eye(4, dtype=np.float32))
    L = _lib.load()
    h = _lib.c_handle()
    t = np.eye(4, dtype=np.float32)
    rc = L.wmd_create(t.ctypes.data_as(_lib.c_f32p), 4, 4, 4, 0, 0, ctypes.byref(h))
    assert rc == -2 and not h.value                               # WMD_ENODEV
    assert b"no CPU fallback" in L.wmd_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "consistent__style_transfer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
