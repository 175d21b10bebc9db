"""Import shim: the product package lives in the directory ``autogp.jl_b200/`` (a name Python
cannot import directly).  ``import autogp.jl_b200`` resolves through this namespace package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "autogp.jl_b200")
_name = __name__ + ".jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _name, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = sys.modules[_name]
