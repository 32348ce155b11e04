"""Import shim: the package directory is `video-style-transfer_b200/` (the name the build
contract fixes); hyphens are not importable, so `import vst_b200` loads that directory
as the package `vst_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "video-style-transfer_b200")
_spec = _u.spec_from_file_location("vst_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["vst_b200"] = _mod
_spec.loader.exec_module(_mod)
