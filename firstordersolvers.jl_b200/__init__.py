"""fos_b200 -- B200-native hot path of mfalt/FirstOrderSolvers.jl behind the package's own API.

The directory is called ``firstordersolvers.jl_b200`` (not importable by that name because of
the dot); ``import fos_b200`` (the loader module at the repository root) registers it.

Exports mirror the reference (src/FirstOrderSolvers.jl:12, src/solvers/solvers.jl:1):
``Feasibility, GAP, GAPA, GAPP, DR, AP, Dykstra, FISTA`` plus the MathProgBase-style methods
of src/FOSSolverInterface.jl with the ``!`` dropped.
"""
from .algorithms import AP, DR, FISTA, GAP, GAPA, GAPP, Dykstra, FOSAlgorithm, LineSearchWrapper
from .model import (AffinePlusLinear, ConeProduct, ConicModel, Feasibility, FeasibilityModel, FeasibilitySolution,
                    FOSMathProgModel, MVHistory, Solution, getobjval, getsolution, loadproblem, numconstr, numvar,
                    optimize, solve, solve_batch, status, supportedcones)
from .model import _Handle as Handle
from ._lib import FosError, lib_path, load as load_library
from . import build as _build
from . import problems

build_library = _build.build

__all__ = ["AP", "DR", "FISTA", "GAP", "GAPA", "GAPP", "Dykstra", "FOSAlgorithm", "LineSearchWrapper", "AffinePlusLinear", "ConeProduct",
           "ConicModel", "Feasibility", "FeasibilityModel", "FeasibilitySolution", "FOSMathProgModel", "MVHistory",
           "Solution", "getobjval", "getsolution", "loadproblem", "numconstr", "numvar", "optimize", "solve", "solve_batch", "status",
           "supportedcones", "Handle", "FosError", "lib_path", "load_library", "build_library", "problems"]
