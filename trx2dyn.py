"""Import shim: the package lives in the directory ``trrosettax2-dynamics_b200/``
(a name Python cannot import directly); ``import trx2dyn`` loads it under this name."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "trrosettax2-dynamics_b200")
_spec = _u.spec_from_file_location("trx2dyn", _os.path.join(_dir, "__init__.py"),
                                   submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["trx2dyn"] = _mod
_spec.loader.exec_module(_mod)
