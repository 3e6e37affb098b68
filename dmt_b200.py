"""Import shim: the package directory is named `diffusionmcmctools.jl_b200` (a dot cannot be written in an import
statement), so load it by path under the module name `diffusionmcmctools_jl_b200` and re-export it here."""
import importlib.util
import os
import sys

_NAME = "diffusionmcmctools_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusionmcmctools.jl_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
pkg = sys.modules[_NAME]
_lib = pkg._lib
configs = pkg.configs
Ctx = pkg.Ctx
DmtError = pkg.DmtError
PKG_DIR = _DIR


def __getattr__(name):
    if name.startswith("__"):  # not a package: `import dmt_b200.x` must not load a second copy of a submodule
        raise AttributeError(name)
    return getattr(pkg, name)
