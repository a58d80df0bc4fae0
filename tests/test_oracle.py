"""CPU tests of the oracle (the checker itself): the reference's two known-answer tests,
an independent O(N^2) numpy evaluation of the pair formulas, and conservation invariants."""
import numpy as np
import pytest

from sphexample_b200 import config, make_params
from sphexample_b200.preprocess import make_particles

import util
from oracle import brute_force


def _default_params(dim=2, **kw):
    consts = config.SimulationConstants(**kw)
    kern = config.SPHKernelInstance(dim, config.WendlandC2(), dx=0.02)
    meta = config.SimulationMetaData(Dimensions=dim)
    return make_params(meta, consts, kern, config.ArtificialViscosity(), config.LinearDensityDiffusion())


def test_time_stepping_kat(oracle_lib):
    """test/runtests.jl:6-16: dt > 0; analytic value 9.03048e-5."""
    p = _default_params()
    parts = make_particles(np.array([[0, 0], [1, 0.0]]), np.array([1000.0, 1000.0]), np.array([1, 1]))
    parts.Acceleration[:] = [[0, 0], [0, -9.81]]
    dt = oracle_lib.Oracle(p, parts).delta_t()
    assert dt > 0
    expected = 0.2 * min(np.sqrt(0.04 / 9.81), 0.04 / (np.sqrt(19.62) * 20))
    assert abs(dt - expected) < 1e-18 + 1e-14 * expected
    assert abs(dt - 9.03048e-5) < 1e-9


def test_isolated_particle_kat(oracle_lib):
    """test/runtests.jl:18-75: one fluid particle in free fall for 1000 iterations of the
    reference's mini-loop: rho == rho0 and P == 0 to 1e-10, x == 0, vx == 0, vz < 0."""
    p = _default_params()
    parts = make_particles(np.array([[0.0, 0.0]]), np.array([1000.0]), np.array([1]))
    o = oracle_lib.Oracle(p, parts)
    o.update_neighbors()
    for _ in range(1000):
        o.neighbor_loop(0)                 # ResetArrays! (no neighbours: zeros)
        dt = o.delta_t()
        o.half_time_step(dt / 2)           # HalfTimeStep + LimitDensityAtBoundary!
        o.pressure(1)
        o.neighbor_loop(1)                 # still no neighbours
        o.full_time_step(dt)               # LimitDensity + DensityEpsi + FullTimeStep
        o.pressure(0)
        assert abs(o.get("rho")[0] - 1000.0) < 1e-10
        assert abs(o.get("press")[0]) < 1e-10
    assert o.get("pos")[0, 0] == 0.0
    assert o.get("vel")[0, 0] == 0.0
    assert o.get("vel")[0, 1] < 0.0


def _subset(case, n_max, lo, hi):
    p = case.particles
    sel = np.all((p.Position >= lo) & (p.Position <= hi), axis=1)
    idx = np.nonzero(sel)[0][:n_max]
    return p.permuted(idx)


@pytest.mark.parametrize("visc,ddt", [(1, 2), (2, 1), (0, 0), (1, 1)])
def test_pair_sums_vs_brute_force_2d(oracle_lib, visc, ddt):
    case = util.perturb(util.case_c1())
    parts = _subset(case, 1500, np.array([0.0, 0.0]), np.array([0.7, 0.5]))
    assert 800 < len(parts) <= 1500
    p = util.params_of(case)
    p.viscosity, p.diffusion = visc, ddt
    p.nu0 = 1e-3
    o = oracle_lib.Oracle(p, parts)
    o.update_neighbors()
    o.pressure(0)
    o.neighbor_loop(0)
    ml = (o.types == 1).astype(np.float64)
    d_bf, a_bf = brute_force.pair_sums(p, o.get("pos"), o.get("rho"), o.get("press"), o.get("vel"), ml,
                                       order_cells=o.cells)
    assert util.relerr(o.get("drhodt"), d_bf) < 1e-12
    assert util.relerr(o.get("acc"), a_bf) < 1e-12


def test_pair_sums_vs_brute_force_3d_pass2(oracle_lib):
    """3D and the pass-2 reads (Q2): positions/velocities/densities of state n+1/2 with the
    diffusion and viscosity terms on the state-n densities."""
    case = util.perturb(util.case_3d_small())
    parts = _subset(case, 1400, np.array([0.0, 0.0, 0.0]), np.array([0.2, 0.22, 0.2]))
    assert 700 < len(parts) <= 1400
    p = util.params_of(case)
    o = oracle_lib.Oracle(p, parts)
    o.update_neighbors()
    o.pressure(0)
    o.neighbor_loop(0)
    o.half_time_step(2e-5)
    o.pressure(1)
    o.neighbor_loop(1)
    ml = (o.types == 1).astype(np.float64)
    d_bf, a_bf = brute_force.pair_sums(p, o.get("pos_h"), o.get("rho_h"), o.get("press"), o.get("vel_h"), ml,
                                       rho_n=o.get("rho"), order_cells=o.cells)
    assert util.relerr(o.get("drhodt"), d_bf) < 1e-12
    assert util.relerr(o.get("acc"), a_bf) < 1e-12


def test_pair_forces_cancel(oracle_lib):
    """sum_i a_i of the pair part is zero to rounding (every pair adds +u and -u)."""
    case = util.perturb(util.case_c1())
    o = oracle_lib.Oracle(util.params_of(case), case.particles)
    o.update_neighbors()
    o.pressure(0)
    o.neighbor_loop(0)
    acc = o.get("acc")
    assert np.max(np.abs(acc.sum(0))) < 1e-9 * np.abs(acc).sum()


def test_cell_list_structure(oracle_lib):
    case = util.case_c1()
    o = oracle_lib.Oracle(util.params_of(case), case.particles)
    ic = o.update_neighbors()
    cells, start = o.cell_list()
    assert ic == len(cells) + 1
    assert start[0] == 0 and start[-1] == len(case.particles)
    # column-major order: last dimension most significant, strictly increasing
    key = cells[:, 1] * 100000 + cells[:, 0]
    assert np.all(np.diff(key) > 0)
    # stable: ids ascending inside every cell on the first sort (input is sorted by ID)
    ids = o.ids
    for c in range(0, len(cells), 37):
        seg = ids[start[c]:start[c + 1]]
        assert np.all(np.diff(seg) > 0)
    assert np.array_equal(o.cells, brute_force.cell_coords(o.get("pos"), util.params_of(case).H_inv))


def test_threads_do_not_change_cell_list_or_state(oracle_lib):
    """the multi-threaded oracle (per-thread accumulators + reduce) agrees with 1 thread to rounding"""
    case = util.perturb(util.case_c1())
    p = util.params_of(case)
    a = oracle_lib.Oracle(p, case.particles, nthreads=1)
    b = oracle_lib.Oracle(p, case.particles, nthreads=4)
    a.step(3, True)
    b.step(3, True)
    assert np.array_equal(a.ids, b.ids)
    assert util.relerr(a.get("rho"), b.get("rho")) < 1e-13
    assert util.relerr(a.get("vel"), b.get("vel")) < 1e-11


@pytest.mark.parametrize("name", ["c1_2d", "3d_small", "c5_mdbc"])
def test_oracle_reproduces_its_committed_outputs(oracle_lib, name):
    """tests/golden/oracle_*.npz (made by tests/golden/make_oracle_golden.py): the restatement's own
    outputs on the shipped layouts at a fixed step, frozen against drift; any thread count must
    reproduce them to rounding"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_oracle_golden", os.path.join(util.GOLDEN, "make_oracle_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(util.GOLDEN, f"oracle_{name}.npz"))
    mk, steps = mod.CASES[name]
    assert steps == int(g["steps"])
    case = mk()
    o = oracle_lib.Oracle(util.params_of(case), case.particles, nthreads=4)
    o.step(steps, True)
    ids = o.ids
    where = {int(i): k for k, i in enumerate(ids)}
    pick = np.array([where[int(i)] for i in g["ids"]])
    assert o.report()["n_rebuilds"] == int(g["n_rebuilds"])
    assert o.report()["total_time"] == pytest.approx(float(g["total_time"]), rel=1e-13)
    assert util.relerr(o.get("rho")[pick], g["rho"]) < 1e-12
    assert util.relerr(o.get("vel")[pick], g["vel"]) < 1e-9
    assert util.relerr(o.get("pos")[pick], g["pos"]) < 1e-13
    assert float(o.get("rho").sum()) == pytest.approx(float(g["sum_rho"]), rel=1e-13)
