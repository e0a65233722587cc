"""Import shim: the package directory is named ``unpaired-multimodal-learning_b200`` (after the
reference repository), which is not a valid Python identifier.  ``import uml_b200`` loads that
directory as the package ``uml_b200`` (submodules resolve normally: ``uml_b200.ops`` ...)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "unpaired-multimodal-learning_b200")
_spec = _ilu.spec_from_file_location("uml_b200", _os.path.join(_dir, "__init__.py"),
                                     submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["uml_b200"] = _mod
_spec.loader.exec_module(_mod)
