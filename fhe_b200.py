"""Import shim: the package directory is named `gpu-homomorphic-encryption_b200` (not a valid Python identifier),
so it is loaded by path and exposed as the module `fhe_b200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu-homomorphic-encryption_b200")
_spec = importlib.util.spec_from_file_location(
    "fhe_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fhe_b200"] = _mod
_spec.loader.exec_module(_mod)
