"""Import helper: the package directory is named `simplemcmc.jl_b200` (a dot is not importable as a plain module
name), so it is registered under the module name `simplemcmc_jl_b200`."""
import importlib.util
import os
import sys

_NAME = "simplemcmc_jl_b200"


def load_package():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    root = os.path.dirname(os.path.abspath(__file__))
    pkg_dir = os.path.join(root, "simplemcmc.jl_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
