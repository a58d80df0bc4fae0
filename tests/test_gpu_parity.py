"""GPU parity tests: libsphb200.so (through the C-ABI) against the CPU oracle on the same inputs.

Tolerances, relative to max|field| (fp64 1e-12 per pass / 1e-8 after 100 steps; fp32
1e-5 per pass / 1e-3 after 100 steps).  The fp32 tolerances below are <= 3x the margin measured on
B200 (profiles/r2l_parity.txt): dρ/dt of one pass 6.6e-7, acceleration 1.3e-5 — slightly ABOVE the
survey's 1e-5 because x_a - x_b of two fp32 positions at |x| ~ 1 m carries an absolute error of
~1e-7 m against pair distances of ~1e-2 m (1e-5 relative); the MUFU approximations of the fast pair
body are 1-2 ulp and do not show.  After 60-100 steps: velocity 8e-6 .. 1.5e-4, density <= 6e-6.
Integer / index work (cell coordinates, cell ranges, the sort permutation) is bit-exact.
"""
import numpy as np
import pytest

import util
from sphexample_b200 import _abi, config
from sphexample_b200.simulation import Simulation, SphError

pytestmark = pytest.mark.gpu

TOL_PASS = {"float64": 1e-12, "float32": 4e-5}      # acceleration; dρ/dt: TOL_PASS_DRHO
TOL_PASS_DRHO = {"float64": 1e-12, "float32": 2e-6}


def make(case, oracle_lib, geometry=(), tweak=None, options=None, nthreads=4):
    p = util.params_of(case, geometry)
    if tweak:
        tweak(p)
    sim = Simulation(p)
    for k, v in (options or {}).items():
        sim.set_option(k, v)
    sim.upload(case.particles)
    orc = oracle_lib.Oracle(p, case.particles, nthreads=nthreads)
    return sim, orc, p


CASES = {
    "c1_2d_f64": lambda: util.perturb(util.case_c1("float64")),
    "c1_2d_f32": lambda: util.perturb(util.case_c1("float32")),
    "3d_f64": lambda: util.perturb(util.case_3d_small("float64")),
    "3d_f32": lambda: util.perturb(util.case_3d_small("float32")),
}


# ------------------------------------------------------------------------------------------------
# cell list (K1, K2): bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_2d_f64", "3d_f64", "3d_f32"])
def test_update_neighbors_matches_oracle(oracle_lib, name):
    case = CASES[name]()
    sim, orc, p = make(case, oracle_lib)
    ic = sim.UpdateNeighbors()
    assert ic == orc.update_neighbors()
    cells, start = sim.cell_list()
    ocells, ostart = orc.cell_list()
    assert np.array_equal(cells, ocells)
    assert np.array_equal(start, ostart)
    st = sim.download()
    assert np.array_equal(st["ID"], orc.ids)                  # same permutation => stable sort
    assert np.array_equal(st["Cells"], orc.cells)
    assert np.array_equal(st["Type"], orc.types)
    assert np.array_equal(st["Position"].astype(np.float64), orc.get("pos"))
    # a second rebuild from the sorted state is the identity
    sim.UpdateNeighbors()
    assert np.array_equal(sim.download(fields=("ID",))["ID"], orc.ids)
    sim.close()


# ------------------------------------------------------------------------------------------------
# one interaction pass (K4), every kernel variant
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("compact,tma", [(1, 1), (0, 1), (1, 0), (0, 0)])
def test_neighbor_loop_pass0(oracle_lib, name, compact, tma):
    case = CASES[name]()
    sim, orc, p = make(case, oracle_lib, options={"compact": compact, "tma": tma})
    sim.UpdateNeighbors()
    orc.update_neighbors()
    sim.Pressure(0)
    orc.pressure(0)
    d, a = sim.NeighborLoop(0)
    orc.neighbor_loop(0)
    util.check(util.relerr(d, orc.get("drhodt")), TOL_PASS_DRHO[case.meta.FloatType])
    util.check(util.relerr(a, orc.get("acc")), TOL_PASS[case.meta.FloatType])
    sim.close()


@pytest.mark.parametrize("name", list(CASES))
def test_staged_step_matches_oracle(oracle_lib, name):
    """the reference's loop body, stage by stage (S2..S18), incl. the pass-2 reads of state n (Q2)"""
    case = CASES[name]()
    sim, orc, p = make(case, oracle_lib)
    f32 = case.meta.FloatType == "float32"
    tol = TOL_PASS[case.meta.FloatType]
    sim.UpdateNeighbors(); orc.update_neighbors()
    sim.Pressure(0); orc.pressure(0)
    sim.NeighborLoop(0); orc.neighbor_loop(0)
    dt = 2.0e-5
    sim.HalfTimeStep(dt / 2); orc.half_time_step(dt / 2)
    half = sim.download_half()
    util.check(util.relerr(half["Position"], orc.get("pos_h")), (1e-7 if f32 else 1e-15))
    util.check(util.relerr(half["Velocity"], orc.get("vel_h")), (1e-6 if f32 else tol))       # measured 2.4e-7
    util.check(util.relerr(half["Density"], orc.get("rho_h")), (1e-7 if f32 else 1e-14))      # measured 3.0e-8
    sim.Pressure(1); orc.pressure(1)
    d, a = sim.NeighborLoop(1)
    orc.neighbor_loop(1)
    util.check(util.relerr(d, orc.get("drhodt")), (4e-6 if f32 else tol))                      # measured 1.3e-6
    util.check(util.relerr(a, orc.get("acc")), (5e-5 if f32 else tol))                         # measured 1.8e-5
    sim.FullTimeStep(dt); orc.full_time_step(dt)
    st = sim.download()
    util.check(util.relerr(st["Position"], orc.get("pos")), (1e-7 if f32 else 1e-15))
    util.check(util.relerr(st["Velocity"], orc.get("vel")), (2e-6 if f32 else tol))            # measured 5.8e-7
    util.check(util.relerr(st["Density"], orc.get("rho")), (5e-7 if f32 else 1e-14))           # measured 1.5e-7
    util.check(util.relerr(st["Acceleration"], orc.get("acc")), (5e-5 if f32 else tol))        # measured 1.8e-5
    sim.close()


@pytest.mark.parametrize("name", ["c1_2d_f64", "3d_f32"])
def test_kernel_variants_are_bitwise_identical(oracle_lib, name):
    """compaction, TMA staging and multi-stage shared memory only change HOW candidates are
    visited, not the order of the per-particle sums => bit-identical results"""
    case = CASES[name]()
    ref = None
    for opts in ({"compact": 1, "tma": 1}, {"compact": 0, "tma": 1}, {"compact": 1, "tma": 0},
                 {"compact": 1, "tma": 1, "smem_kb": 28}, {"compact": 0, "tma": 0, "smem_kb": 20}):
        p = util.params_of(case)
        sim = Simulation(p)
        for k, v in opts.items():
            sim.set_option(k, v)
        sim.upload(case.particles)
        sim.UpdateNeighbors()
        d, a = sim.NeighborLoop(0)
        sim.HalfTimeStep(1e-5)
        d1, a1 = sim.NeighborLoop(1)
        cur = (d, a, d1, a1)
        if ref is None:
            ref = cur
        else:
            for x, y in zip(cur, ref):
                assert np.array_equal(x, y), opts
        sim.close()


# ------------------------------------------------------------------------------------------------
# every model the reference dispatches on, through the generic pair body
# ------------------------------------------------------------------------------------------------
MODELS = {
    "forced_generic": dict(),
    "laminar": dict(viscosity=_abi.VISC_LAMINAR, nu0=1e-3),
    "laminar_sps": dict(viscosity=_abi.VISC_LAMINAR_SPS, nu0=1e-3),
    "zero_zero": dict(viscosity=_abi.VISC_ZERO, diffusion=_abi.DDT_ZERO),
    "zero_gravity_linear": dict(diffusion=_abi.DDT_ZERO_GRAVITY_LINEAR),
    "complex_ddt": dict(diffusion=_abi.DDT_COMPLEX),
    "shifting_kernel_output": dict(shifting=1, kernel_output=1),
}


@pytest.mark.parametrize("name", ["c1_2d_f64", "3d_f64"])
@pytest.mark.parametrize("model", list(MODELS))
def test_model_variants_match_oracle(oracle_lib, name, model):
    case = CASES[name]()

    def tweak(p):
        for k, v in MODELS[model].items():
            setattr(p, k, v)
    sim, orc, p = make(case, oracle_lib, tweak=tweak, options={"generic": 1})
    sim.UpdateNeighbors(); orc.update_neighbors()
    sim.Pressure(0); orc.pressure(0)
    d, a = sim.NeighborLoop(0)
    orc.neighbor_loop(0)
    util.check(util.relerr(d, orc.get("drhodt")), 1e-11)
    util.check(util.relerr(a, orc.get("acc")), 1e-11)
    dt = 2.0e-5
    sim.HalfTimeStep(dt / 2); orc.half_time_step(dt / 2)
    sim.Pressure(1); orc.pressure(1)
    d, a = sim.NeighborLoop(1)
    orc.neighbor_loop(1)
    util.check(util.relerr(d, orc.get("drhodt")), 1e-11)
    util.check(util.relerr(a, orc.get("acc")), 1e-11)
    if p.shifting:
        aux = sim.download_aux()
        util.check(util.relerr(aux["gradC"], orc.get("gradC")), 1e-11)
        util.check(util.relerr(aux["divr"], orc.get("divr")), 1e-11)
        util.check(util.relerr(aux["Kernel"], orc.get("kern")), 1e-11)
        util.check(util.relerr(aux["KernelGradient"], orc.get("kgrad")), 1e-11)
    sim.FullTimeStep(dt); orc.full_time_step(dt)
    st = sim.download()
    util.check(util.relerr(st["Position"], orc.get("pos")), 1e-14)
    util.check(util.relerr(st["Velocity"], orc.get("vel")), 1e-11)
    # and the fused loop on the same model
    sim.step(3); orc.step(3)
    st = sim.download(order="id")
    util.check(util.relerr(st["Velocity"], util.by_id(orc.ids, orc.get("vel"))), 1e-9)
    util.check(util.relerr(st["Density"], util.by_id(orc.ids, orc.get("rho"))), 1e-11)
    sim.close()


# ------------------------------------------------------------------------------------------------
# the fused loop: SimulationLoop semantics, rebuild cadence, Δt
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,nsteps,tol_v,tol_r", [
    ("c1_2d_f64", 1, 1e-12, 1e-14), ("c1_2d_f64", 10, 1e-10, 1e-12), ("c1_2d_f64", 100, 1e-8, 1e-10),
    ("3d_f64", 10, 1e-10, 1e-12), ("3d_f64", 60, 1e-8, 1e-10),
    ("c1_2d_f32", 100, 4e-4, 1.5e-5), ("3d_f32", 60, 2.5e-5, 8e-6),     # measured 1.5e-4 / 5.8e-6 and 8.4e-6 / 2.8e-6
])
def test_fused_steps_match_oracle(oracle_lib, name, nsteps, tol_v, tol_r):
    case = CASES[name]()
    sim, orc, p = make(case, oracle_lib)
    rep = sim.step(nsteps, reset_delta_x=True)
    orc.step(nsteps, True)
    orep = orc.report()
    assert rep["iteration"] == orep["iteration"] == nsteps
    f32 = case.meta.FloatType == "float32"
    assert rep["total_time"] == pytest.approx(orep["total_time"], rel=1e-5 if f32 else 1e-12)
    assert rep["current_dt"] == pytest.approx(orep["current_dt"], rel=1e-4 if f32 else 1e-11)
    if not f32:
        assert rep["n_rebuilds"] == orep["n_rebuilds"]
    st = sim.download(order="id")
    ids = orc.ids
    assert np.array_equal(st["ID"], np.sort(ids))
    util.check(util.relerr(st["Position"], util.by_id(ids, orc.get("pos"))), (1e-6 if f32 else 1e-12))      # measured 3.5e-7
    util.check(util.relerr(st["Velocity"], util.by_id(ids, orc.get("vel"))), tol_v)
    util.check(util.relerr(st["Density"], util.by_id(ids, orc.get("rho"))), tol_r)
    util.check(util.relerr(st["Pressure"], util.by_id(ids, np.asarray(_press(p, orc.get("rho"))))), (2e-3 if f32 else 1e-8))   # measured 6.9e-4 (2D), 7.5e-5 (3D)
    sim.close()


def _press(p, rho):
    return (p.c0 * p.c0 * p.rho0 / 7.0) * ((rho / p.rho0) ** 7 - 1.0)


def test_simulation_loop_matches_oracle(oracle_lib):
    """SimulationLoop: `while TotalTime <= next_output_time`, forced rebuild on entry (Q4)"""
    case = CASES["c1_2d_f64"]()
    sim, orc, p = make(case, oracle_lib)
    for target in (0.0, 2.0e-4, 5.0e-4):
        rep = sim.SimulationLoop(target)
        orc.simulation_loop(target)
        orep = orc.report()
        assert rep["iteration"] == orep["iteration"]
        assert rep["n_rebuilds"] == orep["n_rebuilds"]
        assert rep["total_time"] == pytest.approx(orep["total_time"], rel=1e-12)
        assert rep["total_time"] > target
    full = sim.report()
    assert full["index_counter"] == orep["index_counter"]
    st = sim.download()
    assert np.array_equal(st["ID"], orc.ids)          # identical order after identical rebuilds
    util.check(util.relerr(st["Velocity"], orc.get("vel")), 1e-9)
    util.check(util.relerr(st["Density"], orc.get("rho")), 1e-11)
    sim.close()


def test_delta_t_known_answer_and_oracle(oracle_lib):
    """test/runtests.jl:6-16 on the device: dt = 9.03048e-5"""
    from sphexample_b200 import make_params
    from sphexample_b200.preprocess import make_particles
    consts = config.SimulationConstants()
    kern = config.SPHKernelInstance(2, config.WendlandC2(), dx=0.02)
    meta = config.SimulationMetaData(Dimensions=2)
    p = make_params(meta, consts, kern, config.ArtificialViscosity(), config.LinearDensityDiffusion())
    parts = make_particles(np.array([[0, 0], [1, 0.0]]), np.array([1000.0, 1000.0]), np.array([1, 1]))
    parts.Acceleration[:] = [[0, 0], [0, -9.81]]
    sim = Simulation(p).upload(parts)
    dt = sim.DeltaT()
    assert dt > 0 and abs(dt - 9.03048e-5) < 1e-9
    assert dt == pytest.approx(oracle_lib.Oracle(p, parts).delta_t(), rel=1e-14)
    sim.close()
    case = CASES["3d_f64"]()
    sim, orc, p = make(case, oracle_lib)
    sim.step(3, True); orc.step(3, True)
    assert sim.DeltaT() == pytest.approx(orc.delta_t(), rel=1e-10)
    sim.close()


def test_isolated_particle_known_answer():
    """test/runtests.jl:18-75 on the device, through the fused loop: free fall, rho == rho0, P == 0"""
    from sphexample_b200 import make_params
    from sphexample_b200.preprocess import make_particles
    consts = config.SimulationConstants()
    kern = config.SPHKernelInstance(2, config.WendlandC2(), dx=0.02)
    meta = config.SimulationMetaData(Dimensions=2)
    p = make_params(meta, consts, kern, config.ArtificialViscosity(), config.LinearDensityDiffusion())
    parts = make_particles(np.array([[0.0, 0.0]]), np.array([1000.0]), np.array([1]))
    sim = Simulation(p).upload(parts)
    for _ in range(20):
        sim.step(50)
        st = sim.download()
        assert abs(st["Density"][0] - 1000.0) < 1e-10 and abs(st["Pressure"][0]) < 1e-10
    assert st["Position"][0, 0] == 0.0 and st["Velocity"][0, 0] == 0.0 and st["Velocity"][0, 1] < 0.0
    assert sim.report()["iteration"] == 1000
    sim.close()


# ------------------------------------------------------------------------------------------------
# mDBC (config 5) and moving bodies
# ------------------------------------------------------------------------------------------------
def test_mdbc_matches_oracle(oracle_lib):
    case = util.case_c5("float64")
    assert int(np.any(case.particles.GhostPoints != 0, axis=1).sum()) > 500
    sim, orc, p = make(case, oracle_lib)
    assert p.mdbc == 1
    sim.UpdateNeighbors(); orc.update_neighbors()
    sim.Pressure(0); orc.pressure(0)
    sim.ApplyMDBCBeforeHalf(); orc.apply_mdbc()
    st = sim.download()
    bnd = st["Type"] != 1
    assert np.any(np.abs(orc.get("rho")[bnd] - 1000.0) > 1e-3)      # the correction did something
    util.check(util.relerr(st["Density"], orc.get("rho")), 1e-11)
    d, a = sim.NeighborLoop(0)
    orc.neighbor_loop(0)
    util.check(util.relerr(d, orc.get("drhodt")), 1e-12)
    util.check(util.relerr(a, orc.get("acc")), 1e-12)
    sim.close()
    # the fused loop with S6 inside, 50 steps
    sim, orc, p = make(util.case_c5("float64"), oracle_lib)
    sim.step(50, True); orc.step(50, True)
    st = sim.download(order="id")
    ids = orc.ids
    assert sim.report()["total_time"] == pytest.approx(orc.report()["total_time"], rel=1e-12)
    util.check(util.relerr(st["Density"], util.by_id(ids, orc.get("rho"))), 1e-9)
    util.check(util.relerr(st["Velocity"], util.by_id(ids, orc.get("vel"))), 1e-7)
    sim.close()


def test_progress_motion_matches_oracle(oracle_lib):
    """ProgressMotion (Q8): a block of the wall is re-typed Moving with a MotionDetails entry"""
    case = CASES["c1_2d_f64"]()
    parts = case.particles
    sel = (parts.Type == 2) & (parts.Position[:, 0] < 0.2) & (parts.Position[:, 1] > 0.3)
    assert sel.sum() > 50
    parts.Type[sel] = 3
    parts.GroupMarker[sel] = 9
    geo = [config.Geometry("moving.csv", 9, config.Moving, config.MotionDetails(1.5, 0.0, 1.0, (1.0, 0.0)))]
    sim, orc, p = make(case, oracle_lib, geometry=geo)
    assert p.n_motions == 1
    sim.step(40, True); orc.step(40, True)
    st = sim.download(order="id")
    ids = orc.ids
    mv = st["Type"] == 3
    assert np.all(st["Velocity"][mv][:, 0] == 1.5)
    util.check(util.relerr(st["Position"], util.by_id(ids, orc.get("pos"))), 1e-12)
    util.check(util.relerr(st["Velocity"], util.by_id(ids, orc.get("vel"))), 1e-9)
    util.check(util.relerr(st["Density"], util.by_id(ids, orc.get("rho"))), 1e-11)
    util.check(util.relerr(st["Acceleration"], util.by_id(ids, orc.get("acc"))), 1e-8)
    sim.close()


# ------------------------------------------------------------------------------------------------
# robustness: call order, capacity growth, ragged inputs
# ------------------------------------------------------------------------------------------------
def test_call_order_errors():
    case = CASES["c1_2d_f64"]()
    sim = Simulation(util.params_of(case))
    with pytest.raises(SphError) as e:
        sim.step(1)
    assert e.value.code == _abi.ESTATE
    sim.upload(case.particles)
    with pytest.raises(SphError) as e:
        sim.NeighborLoop(0)
    assert e.value.code == _abi.ESTATE
    sim.close()


def test_cell_grid_grows_when_particles_leave_the_box(oracle_lib):
    """a fluid particle launched far outside the initial bounding box: the dense grid is
    re-allocated (ECAPACITY recovery) and results still match the oracle's unbounded Dict"""
    case = CASES["c1_2d_f64"]()
    parts = case.particles
    k = int(np.nonzero(parts.Type == 1)[0][-1])
    parts.Velocity[k] = [4000.0, 9000.0]
    sim, orc, p = make(case, oracle_lib)
    for _ in range(6):
        sim.step(25, True); orc.step(25, True)
    assert sim.report()["n_rebuilds"] == orc.report()["n_rebuilds"]
    st = sim.download(order="id")
    ids = orc.ids
    assert np.max(st["Position"]) > 15.0
    util.check(util.relerr(st["Position"], util.by_id(ids, orc.get("pos"))), 1e-12)
    util.check(util.relerr(st["Density"], util.by_id(ids, orc.get("rho"))), 1e-10)
    sim.close()


def test_tiny_and_ragged_inputs(oracle_lib):
    """1, 2, 3 and 33 particles (partial warps, empty neighbour rows, negative coordinates)"""
    from sphexample_b200.preprocess import make_particles
    base = util.case_c1()
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 33, 130):
        pos = rng.uniform(-0.15, 0.15, (n, 2))
        parts = make_particles(pos, np.full(n, 1000.0) + rng.uniform(0, 3, n), rng.integers(1, 3, n).astype(np.uint8))
        parts.Velocity[:] = rng.uniform(-1, 1, (n, 2))
        p = util.params_of(base)
        sim = Simulation(p).upload(parts)
        orc = oracle_lib.Oracle(p, parts)
        sim.step(5, True); orc.step(5, True)
        st = sim.download(order="id")
        util.check(util.relerr(st["Velocity"], util.by_id(orc.ids, orc.get("vel"))), 1e-11)
        util.check(util.relerr(st["Density"], util.by_id(orc.ids, orc.get("rho"))), 1e-13)
        sim.close()
