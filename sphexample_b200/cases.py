"""Deterministic lattice generators and parameter sets for the BASELINE configs.

The geometry follows the shipped DualSPHysics exports under input/dam_break_2d and
input/dam_break_3d of the reference (2D: the dp = 0.02 generator reproduces the shipped files'
particle counts exactly, 2 465 + 4 416; 3D: 171 721 vs the shipped 171 496 at dp = 0.0085), so
the same case can be regenerated on a finer lattice: C2 ≈ 60 k (2D), C3 ≈ 1 M, C4 ≈ 16 M (3D).
No RNG anywhere.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import _abi
from .config import (ArtificialViscosity, LinearDensityDiffusion, SimulationConstants,
                     SimulationMetaData, SPHKernelInstance, WendlandC2)
from .preprocess import SimParticles, make_particles


def _r(length: float, dp: float) -> int:
    return int(round(length / dp)) + 1


def hydrostatic_density(depth, rho0, g, c0, gamma=7.0):
    """ρ(depth) from the inverse Tait equation (DualSPHysics' hydrostatic initialisation)."""
    cb = c0 * c0 * rho0 / gamma
    return rho0 * (1.0 + rho0 * g * depth / cb) ** (1.0 / gamma)


def dam_break_2d(dp: float = 0.02, c0: float = 88.14487860902641, dtype=np.float64) -> SimParticles:
    """4 m × 3 m tank, 5-layer U-shaped wall, fluid column 0.9 m × 1.9 m starting 5·dp off the
    corner (input/dam_break_2d/DamBreak2d_Dp0.02_{Bound,Fluid}.csv at dp = 0.02)."""
    nx, nz = _r(4.0, dp), _r(3.0, dp)
    # floor: 5 layers across the full width; sides: 5 layers each, above the floor
    ix, iz = np.meshgrid(np.arange(nx), np.arange(5), indexing="ij")
    floor = np.stack([ix.ravel(), iz.ravel()], 1)
    sx, sz = np.meshgrid(np.concatenate([np.arange(5), np.arange(nx - 5, nx)]), np.arange(5, nz), indexing="ij")
    sides = np.stack([sx.ravel(), sz.ravel()], 1)
    bound = np.concatenate([floor, sides]).astype(np.float64) * dp
    fx, fz = _r(0.9, dp), _r(1.9, dp)
    gx, gz = np.meshgrid(np.arange(fx), np.arange(fz), indexing="ij")
    fluid = (np.stack([gx.ravel(), gz.ravel()], 1).astype(np.float64) + 5.0) * dp
    z_top, z_bot = fluid[:, 1].max(), fluid[:, 1].min()
    depth = (z_top - fluid[:, 1]) * (z_top / max(z_top - z_bot, dp))
    rho = np.concatenate([np.full(len(bound), 1000.0), hydrostatic_density(depth, 1000.0, 9.81, c0)])
    types = np.concatenate([np.full(len(bound), _abi.FIXED), np.full(len(fluid), _abi.FLUID)])
    group = np.concatenate([np.full(len(bound), 1), np.full(len(fluid), 2)])
    return make_particles(np.concatenate([bound, fluid]), rho, types, group, dtype=dtype)


def dam_break_3d(dp: float = 0.0085, c0: float = 33.14, dtype=np.float64) -> SimParticles:
    """SPHERIC-style 3D dam break (input/dam_break_3d/DamBreak3d_Dp0.0085_*): single-layer
    open-top box 1.598 × 0.6715 × 0.3995 m from (0.001,0.001,0.001), a hollow obstacle box with a
    front plate at x ≈ 0.902, fluid lattice 0.391 × 0.6545 × 0.289 m one dp off the corner."""
    o = 0.001
    nx, ny, nz = _r(1.598, dp), _r(0.6715, dp), _r(0.3995, dp)
    # obstacle footprint in lattice indices (plate = its front face)
    pi0 = int(round(0.901 / dp))
    pi1 = pi0 + int(round(0.119 / dp))
    pj0 = int(round(0.238 / dp))
    pj1 = pj0 + int(round(0.119 / dp))
    pk1 = int(round(0.4505 / dp))
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ix, iy = ix.ravel(), iy.ravel()
    inside = (ix >= pi0) & (ix <= pi1) & (iy >= pj0) & (iy <= pj1)
    bottom = np.stack([ix[~inside], iy[~inside], np.zeros((~inside).sum(), np.int64)], 1)
    per = (ix == 0) | (ix == nx - 1) | (iy == 0) | (iy == ny - 1)
    px, py = ix[per], iy[per]
    kk = np.arange(1, nz)
    walls = np.stack([np.repeat(px, len(kk)), np.repeat(py, len(kk)), np.tile(kk, len(px))], 1)
    # obstacle: plate (front face), back face, two y faces, top; open at the bottom
    ox, oy, oz = np.meshgrid(np.arange(pi0, pi1 + 1), np.arange(pj0, pj1 + 1), np.arange(0, pk1 + 1), indexing="ij")
    ox, oy, oz = ox.ravel(), oy.ravel(), oz.ravel()
    shell = (ox == pi0) | (ox == pi1) | (oy == pj0) | (oy == pj1) | (oz == pk1)
    obstacle = np.stack([ox[shell], oy[shell], oz[shell]], 1)
    bound = np.concatenate([bottom, walls, obstacle]).astype(np.float64) * dp + o
    fx, fy, fz = _r(0.391, dp), _r(0.6545, dp), _r(0.289, dp)
    gx, gy, gz = np.meshgrid(np.arange(fx), np.arange(fy), np.arange(fz), indexing="ij")
    fluid = (np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1).astype(np.float64) + 1.0) * dp + o
    depth = fluid[:, 2].max() - fluid[:, 2]
    rho = np.concatenate([np.full(len(bound), 1000.0), hydrostatic_density(depth, 1000.0, 9.81, c0)])
    types = np.concatenate([np.full(len(bound), _abi.FIXED), np.full(len(fluid), _abi.FLUID)])
    group = np.concatenate([np.full(len(bound), 1), np.full(len(fluid), 2)])
    return make_particles(np.concatenate([bound, fluid]), rho, types, group, dtype=dtype)


def dam_break_3d_count(dp: float) -> int:
    """Particle count of dam_break_3d(dp) without building it."""
    nx, ny, nz = _r(1.598, dp), _r(0.6715, dp), _r(0.3995, dp)
    a, b, c = int(round(0.119 / dp)) + 1, int(round(0.119 / dp)) + 1, int(round(0.4505 / dp)) + 1
    bottom = nx * ny - a * b
    walls = 2 * (nx + ny - 2) * (nz - 1)
    obstacle = a * b * c - (a - 2) * (b - 2) * (c - 1)
    fluid = _r(0.391, dp) * _r(0.6545, dp) * _r(0.289, dp)
    return bottom + walls + obstacle + fluid


def dp_for_count_3d(target: int) -> float:
    """Lattice spacing whose 3D dam break has ≈ target particles (bisection on the count)."""
    lo, hi = 5e-4, 0.05
    for _ in range(60):
        mid = math.sqrt(lo * hi)
        if dam_break_3d_count(mid) > target:
            lo = mid
        else:
            hi = mid
    return round(hi, 6)


@dataclass
class Case:
    name: str
    particles: SimParticles
    meta: SimulationMetaData
    consts: SimulationConstants
    kernel: SPHKernelInstance
    viscosity: object
    diffusion: object


def case_dam_break_2d(dp: float = 0.02, float_type: str = "float64", particles: SimParticles = None) -> Case:
    """C1 (dp = 0.02, N = 6 881) / C2 (dp = 0.0058, N = 59 909).  Constants are this repo's choice
    (no upstream script runs the plain 2D pair, SURVEY §3.1 note a): dx = dp, c₀ = 88.1449,
    δᵩ = 0.1, CFL = 0.2, α = 0.01, Wendland k = 2, Artificial + Linear."""
    dtype = np.float64 if float_type == "float64" else np.float32
    parts = particles if particles is not None else dam_break_2d(dp, dtype=dtype)
    consts = SimulationConstants(dx=dp, c0=88.14487860902641, delta_phi=0.1, CFL=0.2, alpha=0.01)
    meta = SimulationMetaData(Dimensions=2, FloatType=float_type, SimulationName="DamBreak2D",
                              SimulationTime=2.0, OutputTimes=0.01)
    kern = SPHKernelInstance(2, WendlandC2(), dx=dp, k=2.0)
    return Case("dam_break_2d", parts.astype(dtype), meta, consts, kern, ArtificialViscosity(), LinearDensityDiffusion())


def case_dam_break_3d(dp: float = 0.0085, float_type: str = "float32", particles: SimParticles = None) -> Case:
    """C3 / C4: constants of example/Dambreak3d.jl:8-15,57-59 with dx = dp, h = √3·dp, k = 2."""
    dtype = np.float64 if float_type == "float64" else np.float32
    parts = particles if particles is not None else dam_break_3d(dp, dtype=dtype)
    consts = SimulationConstants(dx=dp, c0=33.14, alpha=0.1, m0=1000 * dp ** 3, CFL=0.2)
    meta = SimulationMetaData(Dimensions=3, FloatType=float_type, SimulationName="DamBreak3D",
                              SimulationTime=1.6, OutputTimes=0.01)
    kern = SPHKernelInstance(3, WendlandC2(), h=1 * math.sqrt(3 * dp ** 2))
    return Case("dam_break_3d", parts.astype(dtype), meta, consts, kern, ArtificialViscosity(), LinearDensityDiffusion())


def case_still_wedge_mdbc(particles: SimParticles, float_type: str = "float64") -> Case:
    """C5: constants of example/StillWedgeMDBC.jl:7,60,69-70; `particles` comes from the shipped
    StillWedge files (tests/golden/still_wedge_mdbc.npz) with ghost nodes already attached."""
    from .config import SimpleMDBC
    dtype = np.float64 if float_type == "float64" else np.float32
    consts = SimulationConstants(dx=0.02, c0=42.48576250492629, delta_phi=0.1, CFL=0.5)
    meta = SimulationMetaData(Dimensions=2, FloatType=float_type, MDBCMode=SimpleMDBC,
                              SimulationName="StillWedge", SimulationTime=4.0, OutputTimes=0.01)
    kern = SPHKernelInstance(2, WendlandC2(), dx=consts.dx)
    return Case("still_wedge_mdbc", particles.astype(dtype), meta, consts, kern, ArtificialViscosity(), LinearDensityDiffusion())
