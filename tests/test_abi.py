"""CPU tests: the C-ABI library loads and exports every symbol include/aces_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aces_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aces_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built):
    import aces_b200

    lib = aces_b200._lib.load()
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in aces_b200.h but not exported"
    # the Python binding covers the same set
    assert set(aces_b200._lib.SIGNATURES) == set(syms)


def test_version_and_no_cpu_fallback(built):
    import aces_b200

    lib = aces_b200._lib.load()
    assert lib.aces_version() >= 1000
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        # without a GPU the context must fail loudly, not fall back
        with pytest.raises(aces_b200.AcesError):
            aces_b200.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "quantumaces.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "README.md", f"{f} mentions the oracle"
