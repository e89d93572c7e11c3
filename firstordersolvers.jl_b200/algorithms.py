"""Algorithm structs -- same names, positional arguments, defaults and keyword capture as the
reference constructors:

* ``GAP(α=0.8, α1=1.8, α2=1.8; direct=false, kwargs...)``   src/solvers/gap.jl:6-13
* ``DR(α=0.5; kwargs...) = GAP(α, 2.0, 2.0; ...)``           src/solvers/solvers.jl:10
* ``AP(α=1; kwargs...)   = GAP(α, 1.0, 1.0; ...)``           src/solvers/solvers.jl:11
* ``GAPA(α=1.0, β=0.0; direct=false, kwargs...)``            src/solvers/gapa.jl:9-15
* ``FISTA(α=1.0; direct=false, kwargs...)``                  src/solvers/fista.jl:6-11
* ``Dykstra(; direct=false, kwargs...)``                     src/solvers/dykstra.jl:6-10
* ``GAPP(α=0.8, α1=1.8, α2=1.8; direct=true, iproj=100, kwargs...)``  src/solvers/gapproj.jl:6-14

Every other keyword (max_iters, eps, checki, verbose, debug, initx, and unknown keys, which the
reference silently ignores) lands in ``.options`` exactly like ``alg.options``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

ALG_CODES = {"GAP": 0, "GAPA": 1, "FISTA": 2, "Dykstra": 3, "GAPP": 4}


@dataclass
class FOSAlgorithm:
    direct: bool = False
    options: dict = field(default_factory=dict)

    # (code, alpha, alpha1, alpha2, beta, iproj) for fos_set_algorithm
    def _params(self):
        raise NotImplementedError

    def _check_supported(self):
        """direct=true (HSDE.jl:10-15) is served by ``fos_set_direct``: exact projection through the dense
        inverse of I + QQ' built on the device at load time (conic form, m+n+1 <= 16384)."""
        return None


@dataclass
class _GAP(FOSAlgorithm):
    α: float = 0.8
    α1: float = 1.8
    α2: float = 1.8

    def _params(self):
        return (ALG_CODES["GAP"], self.α, self.α1, self.α2, 0.0, 100)


def GAP(α=0.8, α1=1.8, α2=1.8, *, direct=False, **kwargs):
    return _GAP(direct=direct, options=dict(kwargs), α=float(α), α1=float(α1), α2=float(α2))


def DR(α=0.5, **kwargs):
    return GAP(α, 2.0, 2.0, **kwargs)


def AP(α=1, **kwargs):
    return GAP(α, 1.0, 1.0, **kwargs)


@dataclass
class _GAPA(FOSAlgorithm):
    α: float = 1.0
    β: float = 0.0

    def _params(self):
        return (ALG_CODES["GAPA"], self.α, 0.0, 0.0, self.β, 100)


def GAPA(α=1.0, β=0.0, *, direct=False, **kwargs):
    return _GAPA(direct=direct, options=dict(kwargs), α=float(α), β=float(β))


@dataclass
class _FISTA(FOSAlgorithm):
    α: float = 1.0

    def _params(self):
        return (ALG_CODES["FISTA"], self.α, 0.0, 0.0, 0.0, 100)


def FISTA(α=1.0, *, direct=False, **kwargs):
    return _FISTA(direct=direct, options=dict(kwargs), α=float(α))


@dataclass
class _Dykstra(FOSAlgorithm):
    def _params(self):
        return (ALG_CODES["Dykstra"], 0.0, 0.0, 0.0, 0.0, 100)


def Dykstra(*, direct=False, **kwargs):
    return _Dykstra(direct=direct, options=dict(kwargs))


@dataclass
class _GAPP(FOSAlgorithm):
    α: float = 0.8
    α1: float = 1.8
    α2: float = 1.8
    iproj: int = 100

    def _params(self):
        return (ALG_CODES["GAPP"], self.α, self.α1, self.α2, 0.0, int(self.iproj))


def GAPP(α=0.8, α1=1.8, α2=1.8, *, direct=True, iproj=100, **kwargs):
    """Reference default is direct=true (gapproj.jl:14); only direct=False runs on the B200 path."""
    return _GAPP(direct=direct, options=dict(kwargs), α=float(α), α1=float(α1), α2=float(α2), iproj=int(iproj))


@dataclass
class _LineSearchWrapper(FOSAlgorithm):
    """``LineSearchWrapper(alg; lsinterval=100, kwargs...)`` src/wrappers/linesearch.jl:3-24: fields
    ``lsinterval``, ``alg``, ``options`` (= merge(alg.options, kwargs))."""
    lsinterval: int = 100
    alg: FOSAlgorithm = None

    def _params(self):
        return self.alg._params()


def LineSearchWrapper(alg, *, lsinterval=100, **kwargs):
    if not isinstance(alg, (_GAP, _GAPA)):
        # support_linesearch(alg) == Val{:False}: the reference logs an @error and carries on (linesearch.jl:20-22)
        import warnings
        warnings.warn(f"Algorithm {type(alg).__name__} does not support line search")
    opts = dict(alg.options)
    opts.update(kwargs)
    return _LineSearchWrapper(direct=alg.direct, options=opts, lsinterval=int(lsinterval), alg=alg)
