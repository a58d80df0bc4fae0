"""CSV loader semantics of src/PreProcess.jl:12-43,217-243 on synthetic files (the shipped inputs
vary in header quoting and spacing), and a property test of the slab edge planner."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from sphexample_b200 import slab
from sphexample_b200.config import Fixed, Fluid, Geometry
from sphexample_b200.preprocess import AllocateDataStructures, LoadBoundaryNormals, LoadMDBCNormals, LoadSpecificCSV


def _write(path, header, rows):
    with open(path, "w") as fh:
        fh.write(header + "\n")
        for r in rows:
            fh.write(",".join(str(x) for x in r) + "\n")


@pytest.mark.parametrize("header", [
    '"Idp","Vel:0","Vel:1","Vel:2","Rhop","Type","Mk","Points:0","Points:1","Points:2"',
    'Idp, Vel:0, Vel:1, Vel:2, Rhop, Type, Mk, Points:0, Points:1, Points:2',
    '"Idp", "Vel:0", "Vel:1", "Vel:2", "Rhop", "Type", "Mk", "Points:0", "Points:1", "Points:2"',
])
def test_loader_ignores_type_mk_vel_and_shifts_idp(tmp_path, header):
    rows = [(10, 9, 9, 9, 1001.5, 3, 7, 0.1, 0.2, 0.3), (4, 9, 9, 9, 1000.25, 3, 7, 0.4, 0.5, 0.6)]
    p = str(tmp_path / "f.csv")
    _write(p, header, rows)
    pts3, rho, typ, mk, ids = LoadSpecificCSV(3, np.float64, Fluid, 2, p)
    assert pts3.tolist() == [[0.1, 0.2, 0.3], [0.4, 0.5, 0.6]] and rho.tolist() == [1001.5, 1000.25]
    assert typ.tolist() == [int(Fluid)] * 2 and mk.tolist() == [2, 2]        # from the Geometry entry, not the CSV (:36-38)
    assert ids.tolist() == [11, 5]                                            # Idp + 1
    pts2 = LoadSpecificCSV(2, np.float32, Fluid, 2, p)[0]
    assert pts2.dtype == np.float32 and np.allclose(pts2, [[0.1, 0.3], [0.4, 0.6]])   # 2D = columns 0 and 2 (:30-34)


def test_allocate_sorts_by_id_and_derives_factors(tmp_path):
    h = '"Idp","Vel:0","Vel:1","Vel:2","Rhop","Type","Mk","Points:0","Points:1","Points:2"'
    fb, ff = str(tmp_path / "b.csv"), str(tmp_path / "f.csv")
    _write(fb, h, [(0, 0, 0, 0, 1000, 0, 0, 0.0, 0, 0.0), (1, 0, 0, 0, 1000, 0, 0, 0.02, 0, 0.0)])
    _write(ff, h, [(3, 1, 1, 1, 1002, 0, 0, 0.02, 0, 0.04), (2, 1, 1, 1, 1001, 0, 0, 0.0, 0, 0.04)])
    geo = [Geometry(ff, 2, Fluid, None), Geometry(fb, 1, Fixed, None)]          # fluid listed first: the sort restores ID order
    parts = AllocateDataStructures(geo, 2)
    assert parts.ID.tolist() == [1, 2, 3, 4]                                    # sort!(…, by = p -> p.ID), :116
    assert parts.Type.tolist() == [int(Fixed), int(Fixed), int(Fluid), int(Fluid)]
    assert parts.GravityFactor.tolist() == [0, 0, -1, -1] and parts.MotionLimiter.tolist() == [0, 0, 1, 1]   # :78-98
    assert np.all(parts.Velocity == 0) and np.all(parts.Acceleration == 0)      # CSV velocities are ignored (:102-103)
    assert parts.Density.tolist() == [1000, 1000, 1001, 1002]


def test_ghost_nodes_attach_by_row_index(tmp_path):
    h = '"Idp","Vel:0","Vel:1","Vel:2","Rhop","Type","Mk","Points:0","Points:1","Points:2"'
    fb = str(tmp_path / "b.csv")
    _write(fb, h, [(k, 0, 0, 0, 1000, 0, 0, 0.02 * k, 0, 0.0) for k in range(4)])
    parts = AllocateDataStructures([Geometry(fb, 1, Fixed, None)], 2)
    g = str(tmp_path / "g.csv")
    _write(g, '"Points:0","Points:1","Points:2","Normal:0","Normal:1","Normal:2"',
           [(0.02 * k, 0, 0.0, 0.0, 0, 0.03) for k in range(3)])
    pts, ghost, nrm = LoadBoundaryNormals(2, np.float64, g)
    assert np.allclose(ghost, pts + nrm) and np.allclose(nrm[:, 1], 0.03)
    LoadMDBCNormals(parts, g)
    assert np.allclose(parts.GhostPoints[:3, 1], 0.03) and np.all(parts.GhostPoints[3] == 0)   # row index match, rest "no ghost" (Q10)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 8), st.integers(0, 10 ** 6), st.integers(16, 120), st.floats(0.2, 5.0))
def test_plan_edges_properties(world, seed, nlay, skew):
    rng = np.random.default_rng(seed)
    w = rng.random(nlay) ** skew
    coords = rng.choice(nlay, size=4000, p=w / w.sum()) - 7          # layers from -7 up: negative cell coordinates too
    if coords.max() - coords.min() + 1 < 2 * world:
        return
    e = slab.plan_edges(coords, world)
    assert len(e) == world + 1 and e[0] == coords.min() and e[-1] == coords.max() + 1
    assert all(b - a >= 2 for a, b in zip(e, e[1:]))
    own = slab.owner_of(coords, e)
    assert own.min() >= 0 and own.max() <= world - 1
    for r in range(world):
        lo, hi = slab.slab_bounds(e, r)
        sel = coords[own == r]
        assert np.all((sel >= lo) & (sel < hi))
