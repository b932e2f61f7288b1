"""Import helper: the package directory is named after the reference
(`augmentedgplikelihoods.jl_b200`, which is not a valid Python identifier), so it is loaded by path
under the module name `augmentedgplikelihoods_jl_b200`."""
import importlib.util
import os
import sys

NAME = "augmentedgplikelihoods_jl_b200"
PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "augmentedgplikelihoods.jl_b200")


def load_package():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(NAME, None)
        raise
    return mod
