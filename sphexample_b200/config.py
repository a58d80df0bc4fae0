"""Host-side mirror of the reference's configuration structs for the hot path.

Mirrors (names, defaults, assertions) of
  SimulationConstants{T}          src/SimulationConstantsConfiguration.jl:36-52
  SPHKernelInstance{K,D,T}        src/SPHKernels.jl:30-72
  SimulationMetaData{D,T,S,K,B,L} src/SimulationMetaDataConfiguration.jl:28-75
  Geometry / MotionDetails / ParticleType   src/SimulationGeometry.jl:10-30
  the dispatch singletons of src/SPHViscosityModels.jl, src/SPHDensityDiffusionModels.jl

Julia's Unicode field names are accepted as keyword aliases (ρ₀, m₀, α, c₀, γ, δᵩ, ν₀, h⁻¹ …)
so reference scripts translate line by line; attributes use ASCII names.
"""
from __future__ import annotations

import math
import unicodedata
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Optional, Sequence, Union

from . import _abi

_ALIASES = {
    "ρ0": "rho0", "m0": "m0", "α": "alpha", "c0": "c0", "γ": "gamma", "γ⁻¹": "gamma_inv",
    "δᵩ": "delta_phi", "δφ": "delta_phi", "Cb⁻¹": "Cb_inv", "ν0": "nu0", "h⁻¹": "h_inv",
    "H⁻¹": "H_inv", "H²": "H2", "H2": "H2", "αD": "alphaD", "η²": "eta2", "η2": "eta2",
}


def _ascii_kwargs(kwargs):
    out = {}
    for key, val in kwargs.items():
        nk = unicodedata.normalize("NFKC", key)
        nk = _ALIASES.get(key, _ALIASES.get(nk, nk))
        out[nk] = val
    return out


class ParticleType(IntEnum):
    """src/SimulationGeometry.jl:10-14"""
    Fluid = _abi.FLUID
    Fixed = _abi.FIXED
    Moving = _abi.MOVING


Fluid, Fixed, Moving = ParticleType.Fluid, ParticleType.Fixed, ParticleType.Moving


# --- dispatch singletons ----------------------------------------------------------------
class SPHKernel:  # src/SPHKernels.jl:10
    code = -1


class WendlandC2(SPHKernel):
    code = _abi.KERNEL_WENDLANDC2


class CubicSpline(SPHKernel):
    code = _abi.KERNEL_CUBICSPLINE

    def __init__(self, eps: float = 1.0):
        self.eps = eps


class SPHViscosity:  # src/SPHViscosityModels.jl:13
    code = -1


class ZeroViscosity(SPHViscosity):
    code = _abi.VISC_ZERO


class ArtificialViscosity(SPHViscosity):
    code = _abi.VISC_ARTIFICIAL


class Laminar(SPHViscosity):
    code = _abi.VISC_LAMINAR


class LaminarSPS(SPHViscosity):
    code = _abi.VISC_LAMINAR_SPS


class SPHDensityDiffusion:  # src/SPHDensityDiffusionModels.jl:21
    code = -1


class ZeroDensityDiffusion(SPHDensityDiffusion):
    code = _abi.DDT_ZERO


class ZeroGravityLinearDensityDiffusion(SPHDensityDiffusion):
    code = _abi.DDT_ZERO_GRAVITY_LINEAR


class LinearDensityDiffusion(SPHDensityDiffusion):
    code = _abi.DDT_LINEAR


class ComplexDensityDiffusion(SPHDensityDiffusion):
    code = _abi.DDT_COMPLEX


# mode singletons, src/SimulationMetaDataConfiguration.jl:12-26
class NoShifting: code = 0
class PlanarShifting: code = 1
class NoKernelOutput: code = 0
class StoreKernelOutput: code = 1
class NoMDBC: code = 0
class SimpleMDBC: code = 1
class NoLog: code = 0
class StoreLog: code = 1


def _mode_code(mode) -> int:
    return int(mode.code if hasattr(mode, "code") else mode)


class SimulationConstants:
    """SimulationConstants{T}; defaults and asserts as src/SimulationConstantsConfiguration.jl:36-52
    (note the 2D default m₀ = ρ₀·dx²)."""

    def __init__(self, **kwargs):
        kw = _ascii_kwargs(kwargs)
        g = lambda name, default: kw.pop(name, default)
        self.rho0 = float(g("rho0", 1000.0))
        self.dx = float(g("dx", 0.02))
        self.m0 = float(g("m0", self.rho0 * self.dx ** 2))
        self.alpha = float(g("alpha", 0.01))
        self.g = float(g("g", 9.81))
        self.c0 = float(g("c0", math.sqrt(self.g * 2) * 20))
        self.gamma = float(g("gamma", 7.0))
        self.gamma_inv = float(g("gamma_inv", 1.0 / self.gamma))
        self.delta_phi = float(g("delta_phi", 0.1))
        self.CFL = float(g("CFL", 0.2))
        self.Cb = float(g("Cb", (self.c0 ** 2 * self.rho0) / self.gamma))
        self.Cb_inv = float(g("Cb_inv", 1.0 / self.Cb))
        self.nu0 = float(g("nu0", 1e-6))
        self.BlinConstant = float(g("BlinConstant", 0.0066))
        self.SmagorinskyConstant = float(g("SmagorinskyConstant", 0.12))
        if kw:
            raise TypeError(f"unknown SimulationConstants fields: {sorted(kw)}")
        assert self.rho0 > 0, "Density (ρ₀) must be positive"
        assert self.dx > 0, "Grid spacing (dx) must be positive"
        assert self.m0 > 0, "Particle mass (m₀) must be positive"
        assert self.alpha > 0, "Artificial viscosity (α) must be positive"
        assert self.g >= 0, "Gravitational constant (g) must be positive"
        assert self.c0 > 0, "Speed of sound (c₀) must be positive"
        assert self.gamma > 0 and self.gamma_inv > 0
        assert self.delta_phi > 0, "Density variation (δᵩ) must be positive"
        assert self.CFL > 0, "CFL condition (CFL) must be positive"
        assert self.Cb >= 0 and self.Cb_inv >= 0
        assert self.nu0 >= 0, "Kinematic viscosity must be positive"


class SPHKernelInstance:
    """SPHKernelInstance{K,D,T}(kernel; dx | h, k=2), src/SPHKernels.jl:42-72."""

    def __init__(self, dimensions: int, kernel: SPHKernel = None, *, dx=None, h=None, k=2.0):
        kernel = kernel if kernel is not None else WendlandC2()
        if isinstance(kernel, type):
            kernel = kernel()
        if (dx is None) == (h is None):
            raise ValueError("Must provide exactly one of `dx` or `h`")
        self.kernel = kernel
        self.D = int(dimensions)
        self.k = float(k)
        self.h = float(k * dx if dx is not None else h)
        self.h_inv = 1.0 / self.h
        self.H = self.k * self.h
        self.H_inv = 1.0 / self.H
        self.H2 = self.H * self.H
        self.alphaD = self._alphaD()
        self.eta2 = (0.01 * self.h) ** 2
        assert self.k > 0 and self.h > 0 and self.alphaD > 0 and self.eta2 >= 0

    def _alphaD(self) -> float:  # src/SPHKernels.jl:20-27
        h, D = self.h, self.D
        if isinstance(self.kernel, WendlandC2):
            if D == 2:
                return 7 / (4 * math.pi * h ** 2)
            if D == 3:
                return 21 / (16 * math.pi * h ** 3)
            raise ValueError("WendlandC2 has no 1D constant")
        return {1: 2 / (3 * h), 2: 10 / (7 * math.pi * h ** 2), 3: 1 / (math.pi * h ** 3)}[D]


@dataclass
class MotionDetails:
    """src/SimulationGeometry.jl:17-22"""
    Velocity: float
    StartTime: float
    Duration: float
    Direction: Sequence[float]


@dataclass
class Geometry:
    """src/SimulationGeometry.jl:25-30"""
    CSVFile: str
    GroupMarker: int
    Type: ParticleType
    Motion: Optional[MotionDetails] = None


@dataclass
class SimulationMetaData:
    """SimulationMetaData{D,T,SMode,KMode,BMode,LMode}: only the fields the hot path reads or
    writes (src/SimulationMetaDataConfiguration.jl:28-67); UI/output flags are the caller's."""
    Dimensions: int
    FloatType: str = "float64"            # "float64" | "float32"
    ShiftingMode: object = NoShifting
    KernelOutputMode: object = NoKernelOutput
    MDBCMode: object = NoMDBC
    LogMode: object = NoLog
    SimulationName: str = ""
    SaveLocation: str = ""
    Iteration: int = 0
    OutputEach: float = 0.02
    OutputTimes: Union[float, Sequence[float], None] = None
    OutputIterationCounter: int = 0
    StepsTakenForLastOutput: int = 0
    CurrentTimeStep: float = 0.0
    TotalTime: float = 0.0
    SimulationTime: float = 0.0
    IndexCounter: int = 0
    TimeSteps: list = field(default_factory=list)
    ExportSingleVTKHDF: bool = True       # :49 one transient .vtkhdf file (True) or one file per output (False)
    ExportGridCells: bool = False         # :50 also write the occupied cells as an UnstructuredGrid file
    OutputVariables: Optional[Sequence[str]] = None    # :51-65; None = the reference's default list

    def __post_init__(self):
        if self.OutputTimes is None:
            self.OutputTimes = self.OutputEach
        self.FloatType = {"Float64": "float64", "Float32": "float32"}.get(self.FloatType, self.FloatType)
        assert self.FloatType in ("float64", "float32")
        assert self.Dimensions in (2, 3)

    @property
    def real_bytes(self) -> int:
        return 8 if self.FloatType == "float64" else 4


def next_output_time(meta: SimulationMetaData) -> float:
    """src/SPHCellList.jl:687-698"""
    times = meta.OutputTimes
    if isinstance(times, (int, float)):
        return times * meta.OutputIterationCounter
    idx = meta.OutputIterationCounter  # Julia is 1-based: times[idx]
    if idx < 1:
        return 0.0  # (the reference indexes times[0] here, out of bounds under @inbounds)
    if idx < len(times):
        return times[idx - 1]
    return meta.SimulationTime


def make_params(meta: SimulationMetaData, consts: SimulationConstants, kernel: SPHKernelInstance,
                viscosity: SPHViscosity, diffusion: SPHDensityDiffusion,
                geometry: Sequence[Geometry] = ()) -> _abi.Params:
    """Pack the reference's config structs into the C-ABI parameter block."""
    assert kernel.D == meta.Dimensions, "kernel and metadata dimensions differ"
    p = _abi.Params()
    p.abi_version = _abi.ABI_VERSION
    p.dim = meta.Dimensions
    p.real_bytes = meta.real_bytes
    p.kernel = kernel.kernel.code
    p.viscosity = _mode_code(viscosity)
    p.diffusion = _mode_code(diffusion)
    p.shifting = _mode_code(meta.ShiftingMode)
    p.kernel_output = _mode_code(meta.KernelOutputMode)
    p.mdbc = _mode_code(meta.MDBCMode)
    p.rho0, p.dx, p.m0, p.alpha, p.g, p.c0 = consts.rho0, consts.dx, consts.m0, consts.alpha, consts.g, consts.c0
    p.gamma, p.gamma_inv, p.delta_phi, p.cfl = consts.gamma, consts.gamma_inv, consts.delta_phi, consts.CFL
    p.cb, p.cb_inv, p.nu0 = consts.Cb, consts.Cb_inv, consts.nu0
    p.blin_constant, p.smagorinsky_constant = consts.BlinConstant, consts.SmagorinskyConstant
    p.k, p.h, p.h_inv, p.H, p.H_inv, p.H2 = kernel.k, kernel.h, kernel.h_inv, kernel.H, kernel.H_inv, kernel.H2
    p.alphaD, p.eta2 = kernel.alphaD, kernel.eta2
    p.cubic_eps = getattr(kernel.kernel, "eps", 1.0)
    motions = [g for g in geometry if g.Motion is not None]
    if len(motions) > _abi.MAX_MOTIONS:
        raise ValueError(f"at most {_abi.MAX_MOTIONS} moving groups are supported")
    p.n_motions = len(motions)
    for slot, g in enumerate(motions):
        m = p.motions[slot]
        m.group_marker = g.GroupMarker
        m.velocity, m.start_time, m.duration = g.Motion.Velocity, g.Motion.StartTime, g.Motion.Duration
        d = list(g.Motion.Direction) + [0.0] * (3 - len(g.Motion.Direction))
        for k in range(3):
            m.direction[k] = d[k]
    return p
