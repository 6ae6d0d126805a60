"""The C-ABI library builds, loads and exports every symbol include/*.h
declares (no compute calls: there is no GPU on the CPU test tier)."""
import importlib
import os

import pytest


def test_header_symbols_exported(pkg):
    build = importlib.import_module("go-curdleproofs_b200.build")
    build.build()
    lib = pkg.load_library()
    names = pkg.declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert lib.cdl_abi_version() >> 16 == 1


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.CdlError) as ei:
        pkg.Context(0)
    assert ei.value.code == -1  # CDL_ERR_NO_DEVICE


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkgdir = os.path.join(root, "go-curdleproofs_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} references the oracle"
