"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the header
declares, the ctypes table matches the header, and the product path refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "docvision.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dv_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from pdf_table_b200 import build, _lib

    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    names = _declared()
    assert "dv_create" in names and "dv_ctc_greedy" in names and "dv_dbnet_forward" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in docvision.h but not exported"


def test_ctypes_table_matches_header(lib):
    from pdf_table_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()


def test_version(lib):
    assert lib.dv_version() >= 100


def test_no_gpu_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pdf_table_b200 import _lib

    h = ctypes.c_void_p()
    rc = lib.dv_create(b"post", None, 0, 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert "no CPU fallback" in _lib.last_error(None) or "CUDA" in _lib.last_error(None)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pdf_table_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle/"
