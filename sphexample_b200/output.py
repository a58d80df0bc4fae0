"""Output writers (the step after the hot path).

* `write_particles_csv`: particle dump in the DualSPHysics column layout the loader already reads
  (`Idp, Vel:0..2, Rhop, Press, Type, Mk, Points:0..2`, src/PreProcess.jl:12-43) — the parity artefact:
  the state can be diffed against a reference run or fed back as an input case.
* `SaveVTKHDF`, `VTKHDFTransient`, `SetupVTKOutput`: the reference's VTKHDF PolyData files
  (src/ProduceHDFVTK.jl:120-325,461-621), one file per output or one transient file, written through
  the hand-rolled HDF5 writer `hdf5_min` (no libhdf5 / h5py in this image).
* `write_particles_vtp`: the same point data as VTK XML PolyData."""
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


# ---------------------------------------------------------------------------------------------------
# VTKHDF (src/ProduceHDFVTK.jl).  Dataset shapes are HDF5 (row-major) shapes: the reference's Julia
# arrays are column-major, so its 3×N `Points` is the N×3 dataset written here.
# ---------------------------------------------------------------------------------------------------
OUTPUT_VARIABLES = ["ChunkID", "Kernel", "KernelGradient", "Density", "Pressure", "Velocity", "Acceleration", "BoundaryBool",
                    "ID", "Type", "GroupMarker", "GhostPoints", "GhostNormals"]     # SimMetaData.OutputVariables default, :50-64


def to_3d(a: np.ndarray) -> np.ndarray:
    """`to_3d!` (src/AuxiliaryFunctions.jl:28-34): 2D vectors get a zero third component, same element type
    (note: (v1, v2, 0) — not the (x, 0, z) column convention of the input CSVs)."""
    a = np.asarray(a)
    if a.ndim == 1 or a.shape[1] == 3:
        return a
    out = np.zeros((a.shape[0], 3), a.dtype)
    out[:, :a.shape[1]] = a
    return out


def _empty_connectivity(g, name, n_cells_entries):
    c = g.group(name)
    for ds, val in (("NumberOfCells", n_cells_entries), ("NumberOfConnectivityIds", n_cells_entries),
                    ("Connectivity", np.zeros(0, np.int64)), ("Offsets", np.zeros(1, np.int64))):
        c.dataset(ds, np.asarray(val, np.int64))


def SaveVTKHDF(filepath: str, points: np.ndarray, variable_names=(), *args) -> int:
    """`SaveVTKHDF` (src/ProduceHDFVTK.jl:120-160): one static PolyData file — /VTKHDF with Version [2, 3]
    and the ASCII Type attribute, NumberOfPoints, Points, PointData/<name>, Vertices (one vertex cell per
    point) and empty Lines / Polygons / Strips groups.  Returns the file size."""
    from . import hdf5_min as h5
    assert len(variable_names) == len(args), "Same number of variable_names as args is necessary"
    points = to_3d(points)
    n = int(points.shape[0])
    root = h5.Group()
    g = root.group("VTKHDF")
    g.attrs["Version"] = np.array([2, 3], np.int64)
    g.attrs["Type"] = b"PolyData"
    g.dataset("NumberOfPoints", np.array([n], np.int64))
    g.dataset("Points", points)
    pd = g.group("PointData")
    for name, a in zip(variable_names, args):
        pd.dataset(name, to_3d(a))
    v = g.group("Vertices")
    v.dataset("NumberOfCells", np.array([n], np.int64))
    v.dataset("NumberOfConnectivityIds", np.array([n], np.int64))
    v.dataset("Connectivity", np.arange(n, dtype=np.int64))
    v.dataset("Offsets", np.arange(n + 1, dtype=np.int64))
    for name in ("Lines", "Polygons", "Strips"):
        _empty_connectivity(g, name, np.zeros(1, np.int64))
    return h5.write_file(filepath, root)


class VTKHDFTransient:
    """The single-file mode (`GenerateGeometryStructure` + `GenerateStepStructure` + `AppendVTKHDFData`,
    src/ProduceHDFVTK.jl:163-325): every `append` adds one step to /VTKHDF/Steps and to the growing
    NumberOfPoints / Points / PointData datasets.  The reference keeps the HDF5 file open and extends
    chunked datasets; here the rows are spooled and the file — same groups, datasets, shapes, types and
    values, contiguous instead of chunked storage — is laid out by `close()`.

    Reference behaviour kept as it is: Points are Float64 (`fType`), Version is Int32, the four
    connectivity groups get one zero per step in each dataset (no vertex cells in this mode), and
    Steps/NumberOfParts grows by TWO entries per step (:283-285 and :296-298)."""

    def __init__(self, filepath: str, variable_names, *init_args, vertices: bool = False):
        from . import hdf5_min as h5
        assert len(variable_names) == len(init_args), "Same number of variable_names as args is necessary"
        self._h5, self.filepath, self.names, self.vertices = h5, filepath, list(variable_names), bool(vertices)
        self.points = h5.Spool(np.float64, (3,))
        self.vars = []
        for a in init_args:
            a = np.asarray(a)
            self.vars.append(h5.Spool(a.dtype, (3,) if a.ndim == 2 else ()))
        self.values, self.npoints = [], []
        self.conn = h5.Spool(np.int64) if vertices else None
        self.offs = h5.Spool(np.int64) if vertices else None
        self.closed = False

    def append(self, new_step: float, positions: np.ndarray, *args):
        """`AppendVTKHDFData(root, newStep, Positions, variable_names, args...)`"""
        assert len(args) == len(self.vars) and not self.closed
        pos = to_3d(positions)
        n = int(pos.shape[0])
        self.points.append(pos)
        for sp, a in zip(self.vars, args):
            a = to_3d(a)
            assert a.shape[0] == n
            sp.append(a)
        self.values.append(float(new_step))
        self.npoints.append(n)
        if self.vertices:
            self.conn.append(np.arange(n, dtype=np.int64))
            self.offs.append(np.arange(n + 1, dtype=np.int64))

    def close(self) -> int:
        if self.closed:
            return 0
        h5 = self._h5
        ns = len(self.values)
        npts = np.asarray(self.npoints, np.int64)
        poff = np.concatenate([[0], np.cumsum(npts)[:-1]]).astype(np.int64) if ns else np.zeros(0, np.int64)
        root = h5.Group()
        g = root.group("VTKHDF")
        g.attrs["Version"] = np.array([2, 3], np.int32)
        g.attrs["Type"] = b"PolyData"
        g.dataset("NumberOfPoints", npts)
        g.dataset("Points", self.points)
        pd = g.group("PointData")
        for name, sp in zip(self.names, self.vars):
            pd.dataset(name, sp)
        zeros = np.zeros(ns, np.int64)
        cell_off = np.zeros((ns, 4), np.int64)
        conn_off = np.zeros((ns, 4), np.int64)
        for name in ("Vertices", "Lines", "Polygons", "Strips"):
            c = g.group(name)
            if name == "Vertices" and self.vertices:      # option: real vertex cells, so that ParaView renders the points
                c.dataset("NumberOfCells", npts)
                c.dataset("NumberOfConnectivityIds", npts)
                c.dataset("Connectivity", self.conn)
                c.dataset("Offsets", self.offs)
                cell_off[:, 0] = poff
                conn_off[:, 0] = poff
            else:
                for ds in ("NumberOfConnectivityIds", "NumberOfCells", "Offsets", "Connectivity"):
                    c.dataset(ds, zeros)
        st = g.group("Steps")
        st.attrs["NSteps"] = np.int32(ns)
        st.dataset("Values", np.asarray(self.values, np.float64))
        st.dataset("PartOffsets", np.arange(ns, dtype=np.int64))
        st.dataset("NumberOfParts", np.ones(2 * ns, np.int64))
        st.dataset("PointOffsets", poff)
        st.dataset("CellOffsets", cell_off)
        st.dataset("ConnectivityIdOffsets", conn_off)
        pdo = st.group("PointDataOffsets")
        for name in self.names:
            pdo.dataset(name, poff)
        size = h5.write_file(self.filepath, root)
        for sp in [self.points, *self.vars] + ([self.conn, self.offs] if self.vertices else []):
            sp.close()
        self.closed = True
        return size

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def vtk_point_data(state: dict, particles=None, variable_names=None):
    """The `available` dictionary of `save_particle_data` (src/ProduceHDFVTK.jl:563-577) from a
    `Simulation.download()` state (+ the host `SimParticles` table for the columns that never change on
    the device: GhostPoints, GhostNormals, BoundaryBool — matched by ID).  Returns (names, arrays) for
    the requested variables that are available."""
    n = int(np.asarray(state["Position"]).shape[0])
    T = np.asarray(state["Position"]).dtype
    D = int(np.asarray(state["Position"]).shape[1])
    avail = {}
    for k in ("Density", "Pressure", "Velocity", "Acceleration", "ID", "GroupMarker", "Kernel", "KernelGradient"):
        if k in state:
            avail[k] = np.asarray(state[k])
    if "Type" in state:
        typ = np.asarray(state["Type"])
        avail["Type"] = typ.astype(np.int8)                                    # `Int8.(SimParticles.Type)`
        avail["BoundaryBool"] = (typ != 1).astype(np.uint8)                    # src/PreProcess.jl:78-98: 1 - MotionLimiter
    avail.setdefault("ChunkID", np.zeros(n, np.int64))                         # threading detail of the reference's loop; 0 here
    avail.setdefault("Kernel", np.zeros(n, T))
    avail.setdefault("KernelGradient", np.zeros((n, D), T))
    if particles is not None and "ID" in state:
        order = np.argsort(np.asarray(particles.ID), kind="stable")
        row = order[np.searchsorted(np.asarray(particles.ID)[order], np.asarray(state["ID"]))]
        avail["GhostPoints"] = np.asarray(particles.GhostPoints, T)[row]
        avail["GhostNormals"] = np.asarray(particles.GhostNormals, T)[row]
    else:
        avail.setdefault("GhostPoints", np.zeros((n, D), T))
        avail.setdefault("GhostNormals", np.zeros((n, D), T))
    names = [v for v in (OUTPUT_VARIABLES if variable_names is None else variable_names) if v in avail]
    return names, [avail[v] for v in names]


def SetupVTKOutput(save_location: str, simulation_name: str, export_single: bool = True, variable_names=None, particles=None,
                   export_grid_cells: bool = False, H: float = 0.0):
    """`SetupVTKOutput` (src/ProduceHDFVTK.jl:461-621): returns (save_particles(iteration, total_time, state),
    close_files()) — and, with `export_grid_cells` (SimMetaData.ExportGridCells; needs the cell edge H), a third
    function save_grid(iteration, total_time, unique_cells).  Multi-file mode writes `<name>_<iteration, 6 digits>.vtkhdf`
    (grid: `CellGrid_<name>_<iteration>.vtkhdf`), single-file mode appends to `<name>.vtkhdf` (`<name>_GridCells.vtkhdf`)."""
    import os
    base = os.path.join(save_location, simulation_name)
    grid_base = os.path.join(save_location, "CellGrid_" + simulation_name)
    tr = {"w": None, "g": None}

    def save_particles(iteration: int, total_time: float, state: dict):
        names, arrays = vtk_point_data(state, particles, variable_names)
        if not export_single:
            return SaveVTKHDF(f"{base}_{int(iteration):06d}.vtkhdf", state["Position"], names, *arrays)
        if tr["w"] is None:
            tr["w"] = VTKHDFTransient(base + ".vtkhdf", names, *arrays)
        tr["w"].append(total_time, state["Position"], *arrays)

    def save_grid(iteration: int, total_time: float, unique_cells):
        if not export_single:
            return SaveCellGridVTKHDF(f"{grid_base}_{int(iteration):06d}.vtkhdf", H, unique_cells)
        if tr["g"] is None:
            tr["g"] = VTKHDFGridTransient(base + "_GridCells.vtkhdf", H)
        tr["g"].append(total_time, unique_cells)

    def close_files():
        for k in ("w", "g"):
            if tr[k] is not None:
                tr[k].close()

    if export_grid_cells:
        assert H > 0.0, "export_grid_cells needs the cell edge H"
        return save_particles, close_files, save_grid
    return save_particles, close_files


def compute_grid_geometry(H: float, unique_cells: np.ndarray):
    """`compute_grid_geometry` (src/ProduceHDFVTK.jl:38-118): the occupied cells (UniqueCells[n, D], integer cell
    coordinates; cell c is centred on c·H, edge H) as VTK quads (2D, type 9) / hexahedra (3D, type 12).  Returns
    (points[n·2^D, 3], connectivity, offsets, cell_types, cell_data) — cell_data = the 1-based linear index of the
    cell in the bounding box of the occupied cells, x fastest."""
    cells = np.asarray(unique_cells, np.int64)
    n, D = cells.shape
    if D not in (2, 3):
        raise ValueError(f"Dimensionality of UniqueCells must be 2 or 3, got {D}")
    lo = cells.min(axis=0) if n else np.zeros(D, np.int64)
    ext = (cells.max(axis=0) - lo + 1) if n else np.ones(D, np.int64)
    idx = cells - lo
    cell_data = (idx[:, 1] * ext[0] + idx[:, 0] + 1) if D == 2 else ((idx[:, 2] * ext[1] + idx[:, 1]) * ext[0] + idx[:, 0] + 1)
    sx = np.array([-1, 1, 1, -1], np.float64)
    sy = np.array([-1, -1, 1, 1], np.float64)
    c = cells.astype(np.float64) * H
    if D == 2:
        pts = np.zeros((n, 4, 3))
        pts[:, :, 0] = c[:, None, 0] + sx * (H / 2)
        pts[:, :, 1] = c[:, None, 1] + sy * (H / 2)
        vtk_type = 9
    else:
        pts = np.zeros((n, 8, 3))
        pts[:, :, 0] = c[:, None, 0] + np.tile(sx, 2) * (H / 2)
        pts[:, :, 1] = c[:, None, 1] + np.tile(sy, 2) * (H / 2)
        pts[:, :, 2] = c[:, None, 2] + np.repeat([-1.0, 1.0], 4) * (H / 2)
        vtk_type = 12
    nc = pts.shape[1]
    points = pts.reshape(n * nc, 3)
    connectivity = np.arange(n * nc, dtype=np.int64)
    offsets = np.arange(n + 1, dtype=np.int64) * nc
    return points, connectivity, offsets, np.full(n, vtk_type, np.uint8), cell_data.astype(np.int64)


def SaveCellGridVTKHDF(filepath: str, H: float, unique_cells: np.ndarray) -> int:
    """`SaveCellGridVTKHDF` (src/ProduceHDFVTK.jl:416-452): the cell list as a static UnstructuredGrid file.
    `unique_cells` = the occupied cells, e.g. `Simulation.cell_list()[0]`."""
    from . import hdf5_min as h5
    points, connectivity, offsets, cell_types, cell_data = compute_grid_geometry(H, unique_cells)
    root = h5.Group()
    g = root.group("VTKHDF")
    g.attrs["Version"] = np.array([2, 3], np.int64)
    g.attrs["Type"] = b"UnstructuredGrid"
    g.dataset("NumberOfPoints", np.array([points.shape[0]], np.int64))
    g.dataset("NumberOfCells", np.array([cell_types.shape[0]], np.int64))
    g.dataset("NumberOfConnectivityIds", np.array([connectivity.shape[0]], np.int64))
    g.dataset("Points", points)
    g.dataset("Connectivity", connectivity)
    g.dataset("Offsets", offsets)
    g.dataset("Types", cell_types)
    g.group("CellData").dataset("CellData", cell_data)
    g.group("FieldData")
    return h5.write_file(filepath, root)


class VTKHDFGridTransient:
    """The single-file cell-grid output (`GenerateGeometryStructure` / `GenerateStepStructure` with
    vtk_file_type = "UnstructuredGrid" + `AppendVTKHDFGridData`, src/ProduceHDFVTK.jl:163-230,327-414): every
    `append` adds the occupied cells of one output as quads / hexahedra.  As the reference writes it: connectivity
    ids are local to the step (0 .. points-1), `Steps/ConnectivityIdOffsets` and `Steps/PointOffsets` both hold the
    points written before the step, `Steps/CellOffsets` the cells written before it, `CellData/ChunkID` the ChunkID
    column's first n_cells entries (0 here: ChunkID is a threading detail of the reference's loop)."""

    def __init__(self, filepath: str, H: float):
        from . import hdf5_min as h5
        self._h5, self.filepath, self.H = h5, filepath, float(H)
        self.points = h5.Spool(np.float64, (3,))
        self.conn, self.offs, self.cdata, self.chunk = (h5.Spool(np.int64) for _ in range(4))
        self.types = h5.Spool(np.uint8)
        self.values, self.npoints, self.ncells = [], [], []
        self.closed = False

    def append(self, new_step: float, unique_cells: np.ndarray):
        points, connectivity, offsets, cell_types, cell_data = compute_grid_geometry(self.H, unique_cells)
        self.points.append(points)
        self.conn.append(connectivity)
        self.offs.append(offsets)
        self.types.append(cell_types)
        self.cdata.append(cell_data)
        self.chunk.append(np.zeros(len(cell_data), np.int64))
        self.values.append(float(new_step))
        self.npoints.append(int(points.shape[0]))
        self.ncells.append(int(cell_types.shape[0]))

    def close(self) -> int:
        if self.closed:
            return 0
        h5 = self._h5
        ns = len(self.values)
        npts, ncel = np.asarray(self.npoints, np.int64), np.asarray(self.ncells, np.int64)
        before = lambda a: (np.concatenate([[0], np.cumsum(a)[:-1]]).astype(np.int64) if ns else np.zeros(0, np.int64))
        root = h5.Group()
        g = root.group("VTKHDF")
        g.attrs["Version"] = np.array([2, 3], np.int32)
        g.attrs["Type"] = b"UnstructuredGrid"
        g.dataset("NumberOfPoints", npts)
        g.dataset("Points", self.points)
        g.dataset("Connectivity", self.conn)
        g.dataset("NumberOfCells", ncel)
        g.dataset("NumberOfConnectivityIds", npts)
        g.dataset("Offsets", self.offs)
        g.dataset("Types", self.types)
        g.group("FieldData")
        cd = g.group("CellData")
        cd.dataset("CellData", self.cdata)
        cd.dataset("ChunkID", self.chunk)
        st = g.group("Steps")
        st.attrs["NSteps"] = np.int32(ns)
        st.dataset("Values", np.asarray(self.values, np.float64))
        st.dataset("PartOffsets", np.arange(ns, dtype=np.int64))
        st.dataset("NumberOfParts", np.ones(ns, np.int64))
        st.dataset("PointOffsets", before(npts))
        st.dataset("CellOffsets", before(ncel))
        st.dataset("ConnectivityIdOffsets", before(npts))
        st.group("PointDataOffsets")
        size = h5.write_file(self.filepath, root)
        for sp in (self.points, self.conn, self.offs, self.cdata, self.chunk, self.types):
            sp.close()
        self.closed = True
        return size
