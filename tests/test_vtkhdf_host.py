"""VTKHDF output (sphexample_b200/output.py over the hand-rolled HDF5 writer hdf5_min.py) against the file
layout of the reference's writer (src/ProduceHDFVTK.jl:120-325).  No HDF5 library exists in this image,
so the files are read back with tests/h5_minread.py, an independent walk of the on-disk structures that
checks signatures, sizes, alignment, key order and addresses on the way; with h5py installed the same
assertions also run through libhdf5."""
import numpy as np
import pytest

import h5_minread
import util
from sphexample_b200 import hdf5_min, output


def _state(case, shift=0.0):
    p = case.particles
    return {"Position": p.Position + shift, "Velocity": p.Velocity + 0.25, "Acceleration": p.Position * 0 - 9.81,
            "Density": p.Density, "Pressure": p.Density * 0 + 7.5, "ID": p.ID, "Type": p.Type, "GroupMarker": p.GroupMarker}


def _check_static(tree, st, names, arrays):
    n = len(st["ID"])
    a = tree["/VTKHDF@"]
    assert a["Version"].tolist() == [2, 3] and bytes(a["Type"]) == b"PolyData"
    assert tree["/VTKHDF/NumberOfPoints"].tolist() == [n]
    assert np.array_equal(tree["/VTKHDF/Points"], output.to_3d(st["Position"])) and tree["/VTKHDF/Points"].shape == (n, 3)
    for nm, arr in zip(names, arrays):
        got = tree["/VTKHDF/PointData/" + nm]
        assert got.dtype == np.asarray(arr).dtype and np.array_equal(got, output.to_3d(arr)), nm
    v = "/VTKHDF/Vertices/"
    assert tree[v + "NumberOfCells"].tolist() == [n] and tree[v + "NumberOfConnectivityIds"].tolist() == [n]
    assert np.array_equal(tree[v + "Connectivity"], np.arange(n)) and np.array_equal(tree[v + "Offsets"], np.arange(n + 1))
    for g in ("Lines", "Polygons", "Strips"):
        assert tree[f"/VTKHDF/{g}/NumberOfCells"].tolist() == [0] and tree[f"/VTKHDF/{g}/NumberOfConnectivityIds"].tolist() == [0]
        assert tree[f"/VTKHDF/{g}/Connectivity"].shape == (0,) and tree[f"/VTKHDF/{g}/Offsets"].tolist() == [0]


@pytest.mark.parametrize("mk", [lambda: util.case_3d_small("float32"), lambda: util.case_c5("float64")])
def test_static_file_has_the_reference_layout(tmp_path, mk):
    case = mk()
    st = _state(case)
    names, arrays = output.vtk_point_data(st, case.particles)
    assert names == output.OUTPUT_VARIABLES           # all 13 default variables: PointData needs two symbol-table nodes
    path = str(tmp_path / "SimulationName_000003.vtkhdf")
    size = output.SaveVTKHDF(path, st["Position"], names, *arrays)
    rd = h5_minread.Reader(path)
    assert size == len(rd.b)
    tree = rd.tree()
    _check_static(tree, st, names, arrays)
    assert tree["/VTKHDF/PointData/Type"].dtype == np.int8 and tree["/VTKHDF/PointData/BoundaryBool"].dtype == np.uint8
    if case.particles.Position.shape[1] == 2:        # to_3d!: (v1, v2, 0), element type kept
        assert np.all(tree["/VTKHDF/Points"][:, 2] == 0) and tree["/VTKHDF/Points"].dtype == np.float64
        gp = tree["/VTKHDF/PointData/GhostPoints"]
        assert np.any(gp != 0) and np.array_equal(gp[:, :2], case.particles.GhostPoints)


def test_transient_file_appends_steps_like_the_reference(tmp_path):
    case = util.case_c1("float64")
    names = ["Density", "Velocity", "ID"]
    path = str(tmp_path / "Sim.vtkhdf")
    times, states = [0.0, 0.01, 0.025], []
    with output.VTKHDFTransient(path, names, case.particles.Density, case.particles.Velocity, case.particles.ID) as w:
        for k, t in enumerate(times):
            st = _state(case, shift=0.125 * k)
            states.append(st)
            w.append(t, st["Position"], st["Density"] + k, st["Velocity"], st["ID"])
    tree = h5_minread.Reader(path).tree()
    n, ns = len(case.particles), len(times)
    a = tree["/VTKHDF@"]
    assert a["Version"].dtype == np.int32 and a["Version"].tolist() == [2, 3] and bytes(a["Type"]) == b"PolyData"
    assert tree["/VTKHDF/Steps@"]["NSteps"] == ns and tree["/VTKHDF/Steps@"]["NSteps"].dtype == np.int32
    assert tree["/VTKHDF/Steps/Values"].tolist() == times
    assert tree["/VTKHDF/NumberOfPoints"].tolist() == [n] * ns
    assert tree["/VTKHDF/Steps/PointOffsets"].tolist() == [0, n, 2 * n]
    assert tree["/VTKHDF/Steps/PartOffsets"].tolist() == [0, 1, 2]
    assert tree["/VTKHDF/Steps/NumberOfParts"].tolist() == [1] * (2 * ns)         # the reference extends it twice per step
    assert tree["/VTKHDF/Steps/CellOffsets"].shape == (ns, 4) and not tree["/VTKHDF/Steps/CellOffsets"].any()
    assert tree["/VTKHDF/Steps/ConnectivityIdOffsets"].shape == (ns, 4)
    for nm in names:
        assert tree["/VTKHDF/Steps/PointDataOffsets/" + nm].tolist() == [0, n, 2 * n]
    pts = tree["/VTKHDF/Points"]
    assert pts.dtype == np.float64 and pts.shape == (ns * n, 3)
    for k, st in enumerate(states):
        assert np.array_equal(pts[k * n:(k + 1) * n], output.to_3d(st["Position"]))
        assert np.array_equal(tree["/VTKHDF/PointData/Density"][k * n:(k + 1) * n], st["Density"] + k)
        assert np.array_equal(tree["/VTKHDF/PointData/Velocity"][k * n:(k + 1) * n], output.to_3d(st["Velocity"]))
    for g in ("Vertices", "Lines", "Polygons", "Strips"):
        for ds in ("NumberOfCells", "NumberOfConnectivityIds", "Offsets", "Connectivity"):
            assert tree[f"/VTKHDF/{g}/{ds}"].tolist() == [0] * ns
    # option beyond the reference: real vertex cells per step
    p2 = str(tmp_path / "SimV.vtkhdf")
    with output.VTKHDFTransient(p2, ["ID"], case.particles.ID, vertices=True) as w:
        for k, t in enumerate(times[:2]):
            w.append(t, states[k]["Position"], states[k]["ID"])
    t2 = h5_minread.Reader(p2).tree()
    assert t2["/VTKHDF/Vertices/NumberOfCells"].tolist() == [n, n] and t2["/VTKHDF/Vertices/Offsets"].shape == (2 * (n + 1),)
    assert t2["/VTKHDF/Steps/CellOffsets"][:, 0].tolist() == [0, n] and t2["/VTKHDF/Steps/ConnectivityIdOffsets"][:, 0].tolist() == [0, n]


def test_setup_vtk_output_names_files_like_the_reference(tmp_path):
    case = util.case_3d_small("float32")
    st = _state(case)
    save, close = output.SetupVTKOutput(str(tmp_path), "DamBreak", export_single=False, variable_names=["Density", "Velocity"])
    save(0, 0.0, st)
    save(12, 0.12, st)
    close()
    assert sorted(p.name for p in tmp_path.iterdir()) == ["DamBreak_000000.vtkhdf", "DamBreak_000012.vtkhdf"]
    tree = h5_minread.Reader(str(tmp_path / "DamBreak_000012.vtkhdf")).tree()
    _check_static(tree, st, ["Density", "Velocity"], [st["Density"], st["Velocity"]])
    save, close = output.SetupVTKOutput(str(tmp_path), "Single", export_single=True, variable_names=["Density"])
    save(0, 0.0, st)
    save(1, 0.5, st)
    close()
    tree = h5_minread.Reader(str(tmp_path / "Single.vtkhdf")).tree()
    assert tree["/VTKHDF/Steps/Values"].tolist() == [0.0, 0.5] and tree["/VTKHDF/Points"].shape == (2 * len(st["ID"]), 3)


def test_writer_primitives_and_group_fan_out(tmp_path):
    """dtypes, scalar / array / string attributes on groups and datasets, and a group with more links than one
    symbol-table node holds (several nodes under the B-tree, keys in strcmp order)"""
    root = hdf5_min.Group()
    g = root.group("many")
    want = {}
    for k in range(37):
        name = f"d{k:02d}" if k % 3 else f"D_{k}"
        dt = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64][k % 10]
        arr = (np.arange(2 * k + 1) * 3 - 5).astype(dt).reshape(-1, 1) if k % 2 else (np.arange(k) - 4).astype(dt)
        want["/many/" + name] = arr
        d = g.dataset(name, arr)
        if k == 5:
            d.attrs["units"] = "m/s"
            d.attrs["scale"] = np.float32(2.5)
    g.attrs["vec"] = np.array([1.5, -2.0, 4.25])
    root.attrs["note"] = b"root attribute"
    root.group("empty")
    path = str(tmp_path / "t.h5")
    hdf5_min.write_file(path, root)
    tree = h5_minread.Reader(path).tree()
    for k, arr in want.items():
        assert tree[k].dtype == arr.dtype and tree[k].shape == arr.shape and np.array_equal(tree[k], arr), k
    assert tree["/many@"]["vec"].tolist() == [1.5, -2.0, 4.25] and bytes(tree["@"]["note"]) == b"root attribute"
    assert bytes(tree["/many/d05@"]["units"]) == b"m/s" and tree["/many/d05@"]["scale"] == np.float32(2.5)
    with pytest.raises(TypeError):
        hdf5_min.Dataset(np.array(["a", "b"]))


def test_reader_rejects_damaged_files(tmp_path):
    """the checker is only worth something if it notices damage: flip structural bytes and expect a complaint"""
    case = util.case_c1("float64")
    st = _state(case)
    path = str(tmp_path / "ok.vtkhdf")
    output.SaveVTKHDF(path, st["Position"], ["Density"], st["Density"])
    raw = bytearray(open(path, "rb").read())
    h5_minread.Reader(path).tree()
    hits = 0
    for needle in (b"TREE", b"SNOD", b"HEAP"):
        pos = raw.index(needle)
        bad = bytearray(raw)
        bad[pos] ^= 0xFF
        p = str(tmp_path / "bad.h5")
        open(p, "wb").write(bad)
        with pytest.raises(h5_minread.H5Error):
            h5_minread.Reader(p).tree()
        hits += 1
    bad = bytearray(raw[:-8])                          # truncated: the end-of-file address no longer matches
    open(str(tmp_path / "trunc.h5"), "wb").write(bad)
    with pytest.raises(h5_minread.H5Error):
        h5_minread.Reader(str(tmp_path / "trunc.h5"))
    assert hits == 3


def test_cross_check_with_libhdf5_when_available(tmp_path):
    h5py = pytest.importorskip("h5py")
    case = util.case_3d_small("float32")
    st = _state(case)
    names, arrays = output.vtk_point_data(st, case.particles)
    path = str(tmp_path / "x.vtkhdf")
    output.SaveVTKHDF(path, st["Position"], names, *arrays)
    with h5py.File(path, "r") as f:
        g = f["VTKHDF"]
        assert g.attrs["Version"].tolist() == [2, 3] and g.attrs["Type"] == b"PolyData"
        assert np.array_equal(g["Points"][...], st["Position"])
        for nm, arr in zip(names, arrays):
            assert np.array_equal(g["PointData"][nm][...], output.to_3d(arr)), nm
        assert g["Lines/Connectivity"].shape == (0,)


@pytest.mark.parametrize("dim", [2, 3])
def test_cell_grid_file_has_the_reference_layout(tmp_path, dim):
    """SaveCellGridVTKHDF / compute_grid_geometry (src/ProduceHDFVTK.jl:38-118,416-452): one quad / hexahedron per
    occupied cell, centred on cell·H, corner order of the reference, CellData = linear index in the occupied box"""
    from sphexample_b200.slab import cell_coord
    case = util.case_c1("float64") if dim == 2 else util.case_3d_small("float32")
    H = 1.0 / util.params_of(case).H_inv
    cc = np.stack([cell_coord(case.particles.Position[:, k], 1.0 / H) for k in range(dim)], axis=1)
    cells = np.unique(cc, axis=0)
    path = str(tmp_path / "CellGrid_000001.vtkhdf")
    output.SaveCellGridVTKHDF(path, H, cells)
    tree = h5_minread.Reader(path).tree()
    n, nc = len(cells), 2 ** dim
    a = tree["/VTKHDF@"]
    assert a["Version"].tolist() == [2, 3] and bytes(a["Type"]) == b"UnstructuredGrid"
    assert tree["/VTKHDF/NumberOfPoints"].tolist() == [n * nc] and tree["/VTKHDF/NumberOfCells"].tolist() == [n]
    assert tree["/VTKHDF/NumberOfConnectivityIds"].tolist() == [n * nc]
    assert np.array_equal(tree["/VTKHDF/Connectivity"], np.arange(n * nc)) and np.array_equal(tree["/VTKHDF/Offsets"], np.arange(n + 1) * nc)
    assert tree["/VTKHDF/Types"].dtype == np.uint8 and set(tree["/VTKHDF/Types"].tolist()) == {9 if dim == 2 else 12}
    pts = tree["/VTKHDF/Points"].reshape(n, nc, 3)
    assert np.allclose(pts.mean(axis=1)[:, :dim], cells * H, atol=1e-12)                       # centred on cell * H
    assert np.allclose(pts.max(axis=1)[:, :dim] - pts.min(axis=1)[:, :dim], H)                   # edge H
    assert np.allclose(pts[:, 0, :dim], cells * H - H / 2) and np.allclose(pts[:, 2, :2], cells[:, :2] * H + H / 2)   # corner order
    if dim == 2:
        assert np.all(pts[:, :, 2] == 0)
    else:
        assert np.allclose(pts[:, 4, 2], cells[:, 2] * H + H / 2) and np.allclose(pts[:, 3, 2], cells[:, 2] * H - H / 2)
    cd = tree["/VTKHDF/CellData/CellData"]
    lo, ext = cells.min(axis=0), cells.max(axis=0) - cells.min(axis=0) + 1
    k = 5
    want = ((cells[k, 1] - lo[1]) * ext[0] + cells[k, 0] - lo[0] + 1) if dim == 2 else \
        (((cells[k, 2] - lo[2]) * ext[1] + cells[k, 1] - lo[1]) * ext[0] + cells[k, 0] - lo[0] + 1)
    assert cd[k] == want and len(np.unique(cd)) == n and cd.min() >= 1
    # every particle lies inside the box of its own cell
    which = {tuple(c): i for i, c in enumerate(cells)}
    sel = np.arange(0, len(cc), 97)
    for j in sel:
        b = pts[which[tuple(cc[j])]]
        x = case.particles.Position[j].astype(np.float64)
        assert np.all(x >= b.min(axis=0)[:dim] - 1e-9) and np.all(x <= b.max(axis=0)[:dim] + 1e-9)


def test_transient_cell_grid_file_appends_steps_like_the_reference(tmp_path):
    """AppendVTKHDFGridData (src/ProduceHDFVTK.jl:327-414): per step the occupied cells' corners, step-local
    connectivity, and offsets of points / cells written before the step"""
    from sphexample_b200.slab import cell_coord
    case = util.case_c1("float64")
    H = 1.0 / util.params_of(case).H_inv
    steps = []
    for k in range(3):
        x = case.particles.Position + 0.3 * k * H
        steps.append(np.unique(np.stack([cell_coord(x[:, d], 1.0 / H) for d in range(2)], axis=1), axis=0)[: 200 + 17 * k])
    save, close, save_grid = output.SetupVTKOutput(str(tmp_path), "Sim", export_single=True, variable_names=["Density"],
                                                   export_grid_cells=True, H=H)
    for k, cells in enumerate(steps):
        save_grid(k + 1, 0.1 * k, cells)
    close()
    tree = h5_minread.Reader(str(tmp_path / "Sim_GridCells.vtkhdf")).tree()
    nc = np.array([len(c) for c in steps])
    a = tree["/VTKHDF@"]
    assert a["Version"].dtype == np.int32 and bytes(a["Type"]) == b"UnstructuredGrid" and tree["/VTKHDF/Steps@"]["NSteps"] == 3
    assert tree["/VTKHDF/NumberOfCells"].tolist() == nc.tolist() and tree["/VTKHDF/NumberOfPoints"].tolist() == (4 * nc).tolist()
    assert tree["/VTKHDF/NumberOfConnectivityIds"].tolist() == (4 * nc).tolist()
    assert tree["/VTKHDF/Steps/Values"].tolist() == [0.0, 0.1, 0.2]
    assert tree["/VTKHDF/Steps/PointOffsets"].tolist() == [0, 4 * nc[0], 4 * (nc[0] + nc[1])]
    assert tree["/VTKHDF/Steps/ConnectivityIdOffsets"].tolist() == tree["/VTKHDF/Steps/PointOffsets"].tolist()
    assert tree["/VTKHDF/Steps/CellOffsets"].tolist() == [0, nc[0], nc[0] + nc[1]]
    assert tree["/VTKHDF/Steps/PartOffsets"].tolist() == [0, 1, 2] and tree["/VTKHDF/Steps/NumberOfParts"].tolist() == [1, 1, 1]
    assert tree["/VTKHDF/Points"].shape == (4 * nc.sum(), 3) and tree["/VTKHDF/Types"].tolist() == [9] * int(nc.sum())
    assert tree["/VTKHDF/Offsets"].shape == (nc.sum() + 3,)                       # n_cells + 1 per step
    conn = tree["/VTKHDF/Connectivity"]
    assert np.array_equal(conn[4 * nc[0]:4 * (nc[0] + nc[1])], np.arange(4 * nc[1]))   # step-local ids
    p1 = tree["/VTKHDF/Points"][4 * nc[0]:4 * (nc[0] + nc[1])].reshape(nc[1], 4, 3)
    assert np.allclose(p1.mean(axis=1)[:, :2], steps[1] * H)
    assert tree["/VTKHDF/CellData/CellData"].shape == (nc.sum(),) and not tree["/VTKHDF/CellData/ChunkID"].any()
    # multi-file mode: one static grid file per output, named like the reference's
    save, close, save_grid = output.SetupVTKOutput(str(tmp_path), "Multi", export_single=False, export_grid_cells=True, H=H)
    save_grid(7, 0.7, steps[0])
    close()
    t7 = h5_minread.Reader(str(tmp_path / "CellGrid_Multi_000007.vtkhdf")).tree()
    assert t7["/VTKHDF/NumberOfCells"].tolist() == [nc[0]]


def test_random_trees_round_trip(tmp_path):
    """property test of the HDF5 writer: random group trees (depth <= 3, up to 20 links per group, every supported
    dtype, 0-d to 3-d shapes incl. empty ones, attributes) come back identical through the independent reader"""
    from hypothesis import given, settings, strategies as st, HealthCheck
    dtypes = [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]
    names = st.text(alphabet="abcdefghijklmnopqrstuvwxyzABCDEFXYZ_0123456789", min_size=1, max_size=12)
    shapes = st.lists(st.integers(0, 5), min_size=1, max_size=3).map(tuple)

    @st.composite
    def arrays(draw):
        dt = draw(st.sampled_from(dtypes))
        shp = draw(shapes)
        n = int(np.prod(shp))
        seed = draw(st.integers(0, 2 ** 31 - 1))
        a = np.random.default_rng(seed).integers(-100, 100, n).astype(dt).reshape(shp)
        return a

    @st.composite
    def groups(draw, depth=0):
        g = {"attrs": draw(st.dictionaries(names, st.one_of(arrays(), st.binary(min_size=1, max_size=9).filter(lambda b: b"\0" not in b)),
                                           max_size=3)), "kids": {}}
        for nm in draw(st.lists(names, max_size=20 if depth == 0 else 6, unique=True)):
            if depth < 2 and draw(st.integers(0, 4)) == 0:
                g["kids"][nm] = draw(groups(depth + 1))
            else:
                g["kids"][nm] = draw(arrays())
        return g

    def build(spec, node, prefix, want):
        for k, v in spec["attrs"].items():
            node.attrs[k] = v
            want.setdefault(prefix + "@", {})[k] = v
        for nm, v in spec["kids"].items():
            if isinstance(v, dict):
                build(v, node.group(nm), prefix + "/" + nm, want)
            else:
                node.dataset(nm, v)
                want[prefix + "/" + nm] = v

    counter = [0]

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(groups())
    def check(spec):
        root, want = hdf5_min.Group(), {}
        build(spec, root, "", want)
        counter[0] += 1
        path = str(tmp_path / f"r{counter[0]}.h5")
        hdf5_min.write_file(path, root)
        got = h5_minread.Reader(path).tree()
        assert set(got) == set(want)
        for k, v in want.items():
            if k.endswith("@"):
                assert set(got[k]) == set(v)
                for an, av in v.items():
                    if isinstance(av, bytes):
                        assert bytes(got[k][an]) == av
                    else:
                        assert got[k][an].dtype == av.dtype and np.array_equal(got[k][an], av)
            else:
                assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    check()
