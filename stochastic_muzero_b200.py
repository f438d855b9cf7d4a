"""Import name of the package whose sources live in ``stochastic-muzero_b200/`` (a hyphen cannot appear in a
module name): ``import stochastic_muzero_b200`` loads that directory as the package."""
import importlib.util as _util
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "stochastic-muzero_b200")
_spec = _util.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_pkg = _util.module_from_spec(_spec)
_sys.modules[__name__] = _pkg
_spec.loader.exec_module(_pkg)
