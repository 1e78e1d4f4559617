"""Importable alias of the package directory ``4d_plant_analysis_b200`` (a name Python's import
statement cannot spell).  ``import local_icp_b200`` gives the same module object."""
import importlib
import sys

_pkg = importlib.import_module("4d_plant_analysis_b200")
sys.modules[__name__] = _pkg
for _k, _v in list(sys.modules.items()):
    if _k.startswith("4d_plant_analysis_b200."):
        sys.modules["local_icp_b200." + _k.split(".", 1)[1]] = _v
