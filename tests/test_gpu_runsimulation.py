"""RunSimulation (src/SPHCellList.jl:808-930) end to end on the GPU: the outer loop over output intervals
with the reference's own VTKHDF output (transient file and one file per output)."""
import numpy as np
import pytest

import h5_minread
import util
from sphexample_b200.simulation import RunSimulation

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("single", [True, False])
def test_run_simulation_writes_the_reference_vtkhdf_files(tmp_path, single):
    case = util.case_c1("float64")
    m = case.meta
    m.SaveLocation, m.SimulationName, m.ExportSingleVTKHDF = str(tmp_path), "DamBreak2D", single
    m.SimulationTime, m.OutputTimes = 0.004, 0.002
    m.OutputVariables = ["Density", "Pressure", "Velocity", "ID", "Type"]
    m.ExportGridCells = True
    state = RunSimulation(SimMetaData=m, SimConstants=case.consts, SimKernel=case.kernel, SimParticles=case.particles,
                          SimViscosity=case.viscosity, SimDensityDiffusion=case.diffusion)
    n, outs = len(case.particles), m.OutputIterationCounter
    assert outs >= 3 and m.TotalTime > m.SimulationTime and m.Iteration > 10

    def by_id(ids, a):
        return a[np.argsort(ids, kind="stable")]
    if single:
        tree = h5_minread.Reader(str(tmp_path / "DamBreak2D.vtkhdf")).tree()
        assert tree["/VTKHDF/Steps@"]["NSteps"] == outs and tree["/VTKHDF/NumberOfPoints"].tolist() == [n] * outs
        t = tree["/VTKHDF/Steps/Values"]
        assert t[0] == 0.0 and np.all(np.diff(t) > 0) and t[-1] == m.TotalTime
        last = slice((outs - 1) * n, outs * n)
        ids, pts, rho = tree["/VTKHDF/PointData/ID"][last], tree["/VTKHDF/Points"][last], tree["/VTKHDF/PointData/Density"][last]
        first_ids, first_pts = tree["/VTKHDF/PointData/ID"][:n], tree["/VTKHDF/Points"][:n]
        assert np.array_equal(by_id(first_ids, first_pts)[:, :2], by_id(case.particles.ID, case.particles.Position))   # the initial state
        grid = h5_minread.Reader(str(tmp_path / "DamBreak2D_GridCells.vtkhdf")).tree()
        assert grid["/VTKHDF/Steps@"]["NSteps"] == outs and np.all(grid["/VTKHDF/NumberOfCells"] > 400)
        # the particles of the last output lie in the last output's cells
        H, nc = case.kernel.H, grid["/VTKHDF/NumberOfCells"]
        last_cells = grid["/VTKHDF/Points"][-4 * nc[-1]:].reshape(nc[-1], 4, 3).mean(axis=1)[:, :2] / H
        have = {tuple(c) for c in np.rint(last_cells).astype(int)}
        from sphexample_b200.slab import cell_coord
        cc = np.stack([cell_coord(pts[:, d], 1.0 / H) for d in range(2)], axis=1)
        # (the cells are those of the interval's last UpdateNeighbors!: a few surface particles may have left them since)
        assert np.mean([tuple(c) in have for c in cc]) > 0.99
    else:
        names = sorted(p.name for p in tmp_path.iterdir() if not p.name.startswith("CellGrid_"))
        assert names == [f"DamBreak2D_{k:06d}.vtkhdf" for k in range(1, outs + 1)]
        assert sorted(p.name for p in tmp_path.iterdir() if p.name.startswith("CellGrid_")) == \
            [f"CellGrid_DamBreak2D_{k:06d}.vtkhdf" for k in range(1, outs + 1)]
        tree = h5_minread.Reader(str(tmp_path / names[-1])).tree()
        ids, pts, rho = tree["/VTKHDF/PointData/ID"], tree["/VTKHDF/Points"], tree["/VTKHDF/PointData/Density"]
        assert np.array_equal(tree["/VTKHDF/Vertices/Connectivity"], np.arange(n))
    assert np.array_equal(np.sort(ids), state["ID"])
    assert np.array_equal(by_id(ids, pts)[:, :2], state["Position"]) and np.all(pts[:, 2] == 0)
    assert np.array_equal(by_id(ids, rho), state["Density"])
    assert np.abs(state["Velocity"]).max() > 1e-3        # the water has started to move
