"""The C-ABI library loads without a GPU and exports every symbol include/lpfnt.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lpfnt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lpfnt_[a-zA-Z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    from lpfun_b200 import _native

    names = declared_symbols()
    assert len(names) >= 25
    raw = ctypes.CDLL(_native.LIB_PATH)
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing


def test_version_and_error_channel(built_lib):
    from lpfun_b200 import _native as nat

    assert b"sm_100a" in built_lib.lpfnt_version()
    assert built_lib.lpfnt_lp_set_count(0, 3, 2.0) == -1  # LPFNT_EINVAL
    assert "m >= 1" in nat.last_error()
    with pytest.raises(ValueError):
        nat.check(-1)


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "lpfun_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|liblpfnt_oracle|oracle/|orc_[a-z_]+\()")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), os.path.join(dirpath, f)


def test_no_gpu_means_loud_failure(built_lib):
    """Without a CUDA device plan creation must fail with an error, never fall back."""
    from lpfun_b200 import _native as nat

    if nat.device_count() > 0:
        pytest.skip("a GPU is present")
    import lpfun_b200

    with pytest.raises((nat.LpfntError, ValueError)):
        lpfun_b200.Transform(2, 4, report=False)
