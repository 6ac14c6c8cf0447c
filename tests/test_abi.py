"""The C-ABI library loads on a CPU-only box and exports every symbol include/ddpca_b200.h declares."""
import os
import re

import pytest

import ddpca_b200 as dd
from ddpca_b200 import lib as ddlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ddpca_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ddpca_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = dd.load_library()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(ddlib.EXPORTS) == names


def test_abi_version_and_device_count():
    lib = dd.load_library()
    assert lib.ddpca_abi_version() == 2
    assert dd.device_count() >= 0


def test_no_cpu_fallback_without_device():
    if dd.device_count() > 0:
        pytest.skip("a GPU is present")
    from tests.helpers import load_golden

    d, meta, A, P = load_golden("beam_2lev")
    with pytest.raises(dd.DdpcaError, match="no CUDA device"):
        dd.MGPIS.from_hierarchy(A, P)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ddpca-admm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/ref_drivers", "").replace("oracle/_ref", "") or f == "ddpk.py", f
