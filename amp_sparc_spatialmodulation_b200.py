"""Import alias for the package directory ``amp-sparc-spatialmodulation_b200/``.

The directory name required by the repo layout contains hyphens, which Python cannot import
directly.  Importing this module loads that directory as a regular package and registers it in
``sys.modules`` under the importable name ``amp_sparc_spatialmodulation_b200`` (sub-modules such as
``amp_sparc_spatialmodulation_b200.bamp`` resolve through the package ``__path__``).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "amp-sparc-spatialmodulation_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
)
_pkg = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _pkg
_spec.loader.exec_module(_pkg)
