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


def test_cgo_shims_only_use_declared_entry_points(pkg):
    """integration/gpu/*.go (the cgo binding a maintainer drops into the reference) cannot be compiled
    here (no Go toolchain), so at least every C.cdl_* symbol it calls must be declared by the header and
    exported by the built library."""
    import glob
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "integration", "gpu", "*.go")))
    assert len(files) >= 3
    used = set()
    for f in files:
        used |= set(re.findall(r"\bC\.(cdl_[a-z0-9_]+)\s*\(", open(f).read()))
    assert len(used) >= 25
    declared = set(pkg.binding.declared_symbols()) if hasattr(pkg, "binding") else None
    if declared is None:
        import importlib
        declared = set(importlib.import_module("go-curdleproofs_b200.binding").declared_symbols())
    assert used <= declared, sorted(used - declared)
    lib = pkg.load_library()
    for s in used:
        assert hasattr(lib, s), s
