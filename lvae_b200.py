"""Import shim: the product package lives in ``ladder-vae-pytorch_b200/`` (not a valid Python
identifier), this module loads it under the importable name ``lvae_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ladder-vae-pytorch_b200")
_spec = importlib.util.spec_from_file_location(
    "lvae_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["lvae_b200"] = _mod
_spec.loader.exec_module(_mod)
