"""Loader: registers the package directory ``firstordersolvers.jl_b200/`` under the importable
name ``fos_b200`` (a dot cannot appear in a Python module name)."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "firstordersolvers.jl_b200"
_spec = importlib.util.spec_from_file_location("fos_b200", _pkg_dir / "__init__.py",
                                               submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fos_b200"] = _mod
_spec.loader.exec_module(_mod)
