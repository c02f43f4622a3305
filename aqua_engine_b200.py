"""Import shim: the package directory is `aqua-engine_b200/` (hyphen, after the reference
repo's name), which Python cannot import by name.  `import aqua_engine_b200` loads it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "aqua-engine_b200")
_spec = importlib.util.spec_from_file_location(
    "aqua_engine_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["aqua_engine_b200"] = _mod
_spec.loader.exec_module(_mod)
