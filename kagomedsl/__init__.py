"""Import shim: the package directory is literally `kagomedsl.jl_b200/` (the name the project
uses), which Python cannot import by path; this makes `import kagomedsl.jl_b200` resolve to it."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_root = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kagomedsl.jl_b200")
if "kagomedsl.jl_b200" not in _sys.modules:
    _spec = _ilu.spec_from_file_location("kagomedsl.jl_b200", _os.path.join(_root, "__init__.py"),
                                         submodule_search_locations=[_root])
    jl_b200 = _ilu.module_from_spec(_spec)
    _sys.modules["kagomedsl.jl_b200"] = jl_b200
    _spec.loader.exec_module(jl_b200)
else:
    jl_b200 = _sys.modules["kagomedsl.jl_b200"]
