"""Importable alias for the package directory `winmad-s-raytracer-v1.0_b200/` (its name is not a
valid Python identifier):  `import wrt_b200`  loads that package under this name."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "winmad-s-raytracer-v1.0_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
