"""Importable alias for the package directory `multimodal-lip-sync-deepfake-detection-system_b200/`.

The directory name required by the repo layout contains hyphens and cannot be imported with a plain
`import` statement, so this shim loads it under the module name `lipsync_b200` (sub-modules resolve
through `submodule_search_locations`, e.g. `lipsync_b200.model`, `lipsync_b200.inference`).
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.normpath(
    _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..",
                  "multimodal-lip-sync-deepfake-detection-system_b200"))
_spec = _ilu.spec_from_file_location(
    "lipsync_b200", _os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["lipsync_b200"] = _mod
_spec.loader.exec_module(_mod)
