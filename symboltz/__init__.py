"""Import shim: the product package lives in the directory `symboltz.jl_b200/` (a dotted directory name cannot be
imported directly), so this tiny package registers it as the submodule `symboltz.jl_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "symboltz.jl_b200")
if "symboltz.jl_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location("symboltz.jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["symboltz.jl_b200"] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = sys.modules["symboltz.jl_b200"]
