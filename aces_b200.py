"""Loader: makes the package directory `quantumaces.jl_b200/` importable as `aces_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quantumaces.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "aces_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["aces_b200"] = _mod
_spec.loader.exec_module(_mod)

# Input generators (tests / bench.py only) live outside the product package, in inputs/; they are
# attached here so that `aces_b200.rotated_design(9)` keeps working for the test-suite.
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
from inputs import circuit_design as _cd  # noqa: E402

_mod.circuit_design = _cd
_mod.rotated_design, _mod.unrotated_design, _mod.example_design = _cd.rotated_design, _cd.unrotated_design, _cd.example_design
