"""Import shim: the package directory is named cuda-photon-mapper_b200/ (not an importable identifier), so
`import pmb200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda-photon-mapper_b200")
_spec = importlib.util.spec_from_file_location("pmb200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pmb200"] = _mod
_spec.loader.exec_module(_mod)
