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


def write_particles_vtp(path: str, state: dict, select=None) -> int:
    """Write a `Simulation.download()` state as a VTK XML PolyData file (.vtp, binary "appended raw"
    data) that ParaView opens directly: one vertex per particle, point data Velocity, Acceleration
    (when present), Density, Pressure, ID, Type, GroupMarker — the variables the reference's VTKHDF
    writer exports by default (src/ProduceHDFVTK.jl:461-621).  This is NOT the reference's VTKHDF
    container (no HDF5 library is available in this image); it is the same data in the other file
    format ParaView reads natively.  Returns the number of points."""
    import struct
    sel = slice(None) if select is None else select
    pos = _xyz(state["Position"][sel]).astype("<f4")
    n = int(pos.shape[0])
    arrays = [("Velocity", _xyz(state["Velocity"][sel]).astype("<f4"), "Float32", 3)]
    if "Acceleration" in state:
        arrays.append(("Acceleration", _xyz(state["Acceleration"][sel]).astype("<f4"), "Float32", 3))
    arrays.append(("Density", np.asarray(state["Density"][sel]).astype("<f4"), "Float32", 1))
    if "Pressure" in state:
        arrays.append(("Pressure", np.asarray(state["Pressure"][sel]).astype("<f4"), "Float32", 1))
    for name, typ, vtk in (("ID", "<i8", "Int64"), ("Type", "u1", "UInt8"), ("GroupMarker", "<u8", "UInt64")):
        if name in state:
            arrays.append((name, np.asarray(state[name][sel]).astype(typ), vtk, 1))
    conn = np.arange(n, dtype="<i8")
    offs = np.arange(1, n + 1, dtype="<i8")
    blobs, offset = [], 0

    def add(a):
        nonlocal offset
        raw = np.ascontiguousarray(a).tobytes()
        blobs.append(struct.pack("<Q", len(raw)) + raw)
        o = offset
        offset += 8 + len(raw)
        return o
    xml = ['<?xml version="1.0"?>',
           '<VTKFile type="PolyData" version="1.0" byte_order="LittleEndian" header_type="UInt64">', "<PolyData>",
           f'<Piece NumberOfPoints="{n}" NumberOfVerts="{n}" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="0">',
           "<PointData>"]
    for name, a, vtk, nc in arrays:
        xml.append(f'<DataArray type="{vtk}" Name="{name}" NumberOfComponents="{nc}" format="appended" offset="{add(a)}"/>')
    xml += ["</PointData>", "<Points>",
            f'<DataArray type="Float32" Name="Points" NumberOfComponents="3" format="appended" offset="{add(pos)}"/>', "</Points>",
            "<Verts>", f'<DataArray type="Int64" Name="connectivity" format="appended" offset="{add(conn)}"/>',
            f'<DataArray type="Int64" Name="offsets" format="appended" offset="{add(offs)}"/>', "</Verts>", "</Piece>", "</PolyData>",
            '<AppendedData encoding="raw">']
    with open(path, "wb") as fh:
        fh.write(("\n".join(xml) + "\n_").encode())
        for b in blobs:
            fh.write(b)
        fh.write(b"\n</AppendedData>\n</VTKFile>\n")
    return n
