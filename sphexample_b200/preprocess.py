"""Particle table + CSV loader with the reference's semantics (src/PreProcess.jl).

This is the caller side of the hot path: it produces host arrays in exactly
the layout the C-ABI's sphb200_upload expects.
"""
from __future__ import annotations

import csv
from dataclasses import dataclass, fields
from typing import Optional, Sequence

import numpy as np

from . import _abi
from .config import Geometry, ParticleType


@dataclass
class SimParticles:
    """The fields of the reference's SimParticles StructArray that cross the C-ABI
    (src/PreProcess.jl:102-116; canonical list test/runtests.jl:43-48).  Vector fields are
    [N, D] C-contiguous, i.e. the memory layout of Julia's Vector{SVector{D,T}}."""
    Position: np.ndarray
    Velocity: np.ndarray
    Acceleration: np.ndarray
    Density: np.ndarray
    Pressure: np.ndarray
    GravityFactor: np.ndarray
    MotionLimiter: np.ndarray
    BoundaryBool: np.ndarray
    ID: np.ndarray
    Type: np.ndarray
    GroupMarker: np.ndarray
    GhostPoints: np.ndarray
    GhostNormals: np.ndarray
    Cells: Optional[np.ndarray] = None

    def __len__(self):
        return self.Position.shape[0]

    @property
    def Dimensions(self):
        return self.Position.shape[1]

    def astype(self, dtype) -> "SimParticles":
        out = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None and v.dtype.kind == "f":
                v = np.ascontiguousarray(v, dtype=dtype)
            out[f.name] = v
        return SimParticles(**out)

    def permuted(self, perm) -> "SimParticles":
        return SimParticles(**{f.name: (None if getattr(self, f.name) is None else
                                        np.ascontiguousarray(getattr(self, f.name)[perm]))
                               for f in fields(self)})


def gravity_factor_and_motion_limiter(types: np.ndarray, dtype=np.float64):
    """src/PreProcess.jl:78-98 (Q6): GravityFactor Fluid −1 / Moving +1 / Fixed 0;
    MotionLimiter Fluid 1, everything else 0."""
    gf = np.where(types == _abi.FLUID, -1.0, np.where(types == _abi.MOVING, 1.0, 0.0)).astype(dtype)
    ml = np.where(types == _abi.FLUID, 1.0, 0.0).astype(dtype)
    return gf, ml


def make_particles(position, density, types, group_marker=None, ids=None, velocity=None,
                   dtype=np.float64, sort_by_id=True) -> SimParticles:
    """Build the particle table from raw arrays, as AllocateDataStructures does after loading
    (zero velocity/acceleration, derived GravityFactor/MotionLimiter, sort by ID)."""
    position = np.ascontiguousarray(position, dtype=dtype)
    n, d = position.shape
    types = np.ascontiguousarray(types, dtype=np.uint8)
    gf, ml = gravity_factor_and_motion_limiter(types, dtype)
    ids = np.arange(1, n + 1, dtype=np.int64) if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
    group = np.ones(n, dtype=np.uint64) if group_marker is None else np.ascontiguousarray(group_marker, dtype=np.uint64)
    p = SimParticles(
        Position=position,
        Velocity=np.zeros((n, d), dtype) if velocity is None else np.ascontiguousarray(velocity, dtype=dtype),
        Acceleration=np.zeros((n, d), dtype),
        Density=np.ascontiguousarray(density, dtype=dtype),
        Pressure=np.zeros(n, dtype),
        GravityFactor=gf,
        MotionLimiter=ml,
        BoundaryBool=(ml == 0).astype(np.uint8),
        ID=ids,
        Type=types,
        GroupMarker=group,
        GhostPoints=np.zeros((n, d), dtype),
        GhostNormals=np.zeros((n, d), dtype),
        Cells=np.zeros((n, d), np.int64),
    )
    if sort_by_id:
        p = p.permuted(np.argsort(p.ID, kind="stable"))  # sort!(SimParticles, by = p -> p.ID), :116
    return p


def _read_csv_columns(path: str):
    """DualSPHysics/ParaView CSV export; header quoting and spacing vary across the shipped files."""
    with open(path, newline="") as fh:
        reader = csv.reader(fh, skipinitialspace=True)
        header = [h.strip().strip('"').strip() for h in next(reader)]
        rows = [r for r in reader if r]
    cols = {}
    arr = np.array(rows, dtype=object)
    for k, name in enumerate(header):
        col = arr[:, k]
        try:
            cols[name] = np.array([float(x) if x != "" else np.nan for x in col], dtype=np.float64)
        except ValueError:
            cols[name] = col
    return cols


def LoadSpecificCSV(dimensions: int, dtype, particle_type: ParticleType, group_marker: int, path: str):
    """src/PreProcess.jl:12-43: Points:0/1/2 (2D uses columns 0 and 2), Rhop, Idp+1; the CSV's own
    Type / Mk / Vel / Press columns are ignored."""
    c = _read_csv_columns(path)
    p1, p2, p3 = c["Points:0"], c["Points:1"], c["Points:2"]
    pts = np.stack([p1, p2, p3], 1) if dimensions == 3 else np.stack([p1, p3], 1)
    n = pts.shape[0]
    return (pts.astype(dtype), c["Rhop"].astype(dtype), np.full(n, int(particle_type), np.uint8),
            np.full(n, group_marker, np.uint64), c["Idp"].astype(np.int64) + 1)


def AllocateDataStructures(SimGeometry: Sequence[Geometry], dimensions: int, dtype=np.float64) -> SimParticles:
    """src/PreProcess.jl:45-119"""
    parts = [LoadSpecificCSV(dimensions, dtype, g.Type, g.GroupMarker, g.CSVFile) for g in SimGeometry]
    cat = lambda k: np.concatenate([p[k] for p in parts])
    return make_particles(cat(0), cat(1), cat(2), cat(3), cat(4), dtype=dtype)


def LoadBoundaryNormals(dimensions: int, dtype, path: str):
    """src/PreProcess.jl:217-243 → (points, ghost_points = points + normal, normals)"""
    c = _read_csv_columns(path)
    if dimensions == 3:
        normals = np.stack([c["Normal:0"], c["Normal:1"], c["Normal:2"]], 1)
        points = np.stack([c["Points:0"], c["Points:1"], c["Points:2"]], 1)
    else:
        normals = np.stack([c["Normal:0"], c["Normal:2"]], 1)
        points = np.stack([c["Points:0"], c["Points:2"]], 1)
    return points.astype(dtype), (points + normals).astype(dtype), normals.astype(dtype)


def LoadMDBCNormals(particles: SimParticles, path: Optional[str]) -> None:
    """src/SPHCellList.jl:512-524: ghost rows are matched to particles BY ROW INDEX (Q10)."""
    if path is None:
        return
    _, ghost_points, normals = LoadBoundaryNormals(particles.Dimensions, particles.Position.dtype, path)
    n = ghost_points.shape[0]
    particles.GhostPoints[:n] = ghost_points
    particles.GhostNormals[:n] = normals
