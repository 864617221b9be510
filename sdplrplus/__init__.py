"""Import shim: makes the package directory `sdplrplus.jl_b200/` (whose name
contains a dot) importable as `sdplrplus.jl_b200`."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sdplrplus.jl_b200")
_name = __name__ + ".jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_pkg_dir, "__init__.py"),
                                                   submodule_search_locations=[_pkg_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = sys.modules[_name]
