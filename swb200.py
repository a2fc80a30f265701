"""Import shim: the package directory is named `seismicwaves.jl_b200` (not a valid dotted module name),
so load it by path and expose it as `swb200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "seismicwaves.jl_b200")
_spec = importlib.util.spec_from_file_location("swb200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["swb200"] = _mod
_spec.loader.exec_module(_mod)
