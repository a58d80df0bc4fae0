"""Particle dump in the layout of the reference's input CSVs (the step after the
hot path).  The reference writes VTKHDF (src/ProduceHDFVTK.jl), which needs libhdf5 — absent in
this image; a dump in the DualSPHysics column layout the loader already reads
(`Idp, Vel:0..2, Rhop, Press, Type, Mk, Points:0..2`, src/PreProcess.jl:12-43) is the parity
artefact instead: the state can be diffed against a reference run or fed back as an input case."""
from __future__ import annotations

import csv

import numpy as np

HEADER = ["Idp", "Vel:0", "Vel:1", "Vel:2", "Rhop", "Press", "Type", "Mk", "Points:0", "Points:1", "Points:2"]


def _xyz(a: np.ndarray) -> np.ndarray:
    """[N, D] -> [N, 3]; a 2D case lives in the (x, z) columns 0 and 2 (src/PreProcess.jl:30-34)."""
    a = np.asarray(a, np.float64)
    if a.shape[1] == 3:
        return a
    out = np.zeros((a.shape[0], 3))
    out[:, 0], out[:, 2] = a[:, 0], a[:, 1]
    return out


def write_particles_csv(path: str, state: dict, select=None) -> int:
    """Write a `Simulation.download()` state (Position, Velocity, Density, Pressure, ID, Type,
    GroupMarker).  `Idp` is ID − 1, undoing the loader's `Idp + 1`.  `select` = optional boolean mask
    / index array (e.g. one particle type).  Values are written with repr-exact precision.  Returns
    the number of rows."""
    sel = slice(None) if select is None else select
    pos, vel = _xyz(state["Position"][sel]), _xyz(state["Velocity"][sel])
    rho = np.asarray(state["Density"][sel], np.float64)
    prs = np.asarray(state.get("Pressure", np.zeros_like(state["Density"]))[sel], np.float64)
    idp = np.asarray(state["ID"][sel], np.int64) - 1
    typ = np.asarray(state.get("Type", np.zeros(len(idp), np.uint8))[sel], np.int64)
    mk = np.asarray(state.get("GroupMarker", np.zeros(len(idp), np.uint64))[sel], np.int64)
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh, quoting=csv.QUOTE_NONNUMERIC)
        w.writerow(HEADER)
        for k in range(len(idp)):
            w.writerow([int(idp[k]), *map(float, vel[k]), float(rho[k]), float(prs[k]), int(typ[k]), int(mk[k]), *map(float, pos[k])])
    return int(len(idp))
