"""Import shim: the package directory is named `orthogonalsphericalshellgrids.jl_b200` (the
layout the build contract names), which is not a valid Python identifier, so it is registered
here under the importable name `ossg_b200`:

    import ossg_b200 as ossg
    grid = ossg.TripolarGrid(ossg.GPU(), size=(360, 180, 50))
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "orthogonalsphericalshellgrids.jl_b200")
_NAME = "ossg_b200"

_spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod          # replaces this shim so that relative imports resolve
_spec.loader.exec_module(_mod)
