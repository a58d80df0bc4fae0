"""Multi-GPU slab mode through the C-ABI.  With one GPU: world = 1 slab mode (communicator of one,
same code path minus the neighbours) must reproduce the plain single-GPU run bit for bit.  With
>= 2 GPUs: scripts/slab_parity.py under torchrun (slab run vs single-GPU run vs oracle)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from sphexample_b200 import slab
from sphexample_b200.simulation import Simulation

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_of_one_matches_plain_run_bitwise():
    case = util.perturb(util.case_3d_small("float32"))
    p = util.params_of(case)
    ref = Simulation(p)
    ref.upload(case.particles)
    ref.step(40, reset_delta_x=True)
    a = ref.download(order="id")
    ref.close()
    sim = Simulation(p)
    dec = slab.SlabDecomposition(sim, case.particles, p.H_inv, 0, 1, axis=2).setup()
    rep = sim.step(40, reset_delta_x=True)
    b = dec.gather(order="id", fields=("Position", "Velocity", "Density", "Pressure", "ID"))
    sim.close()
    assert rep["n_particles"] == len(case.particles) and rep["n_halo"] == 0
    for k in ("Position", "Velocity", "Density", "Pressure"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name,axis", [("3d_f64", 0), ("3d_f64", 1), ("3d_f64", 2), ("c1_2d_f64", 0), ("c1_2d_f64", 1), ("3d_f32", 0)])
def test_every_slab_axis_reproduces_the_reference_roles(oracle_lib, name, axis):
    """the cell key can have any axis as its slowest component (the slab axis; with x as the slab axis the
    rows run along y): the pair set, the density-diffusion roles (Q1: the reference's cell order, last
    component most significant) and the sort order must come out the same — world of one, against the
    oracle and the plain run"""
    mk = {"3d_f64": lambda: util.case_3d_small("float64"), "3d_f32": lambda: util.case_3d_small("float32"),
          "c1_2d_f64": lambda: util.case_c1("float64")}[name]
    case = util.perturb(mk(), vel_scale=2.0)
    p = util.params_of(case)
    ref = Simulation(p)
    ref.upload(case.particles)
    r0 = ref.step(60, reset_delta_x=True)
    a = ref.download(order="id")
    ref.close()
    sim = Simulation(p)
    dec = slab.SlabDecomposition(sim, case.particles, p.H_inv, 0, 1, axis=axis).setup()
    r1 = sim.step(60, reset_delta_x=True)
    b = dec.gather(order="id", fields=("Position", "Velocity", "Density", "Pressure", "ID"))
    sim.close()
    o = oracle_lib.Oracle(p, case.particles, nthreads=4)
    o.step(60, True)
    f32 = name.endswith("f32")
    assert r0["n_rebuilds"] == r1["n_rebuilds"] >= 2 and (f32 or r1["n_rebuilds"] == o.report()["n_rebuilds"])
    for k, of, tol in (("Position", "pos", 1e-6 if f32 else 1e-13), ("Velocity", "vel", 2.5e-5 if f32 else 1e-9),
                       ("Density", "rho", 8e-6 if f32 else 1e-11)):
        util.check(util.relerr(b[k], a[k]), tol)
        util.check(util.relerr(b[k], util.by_id(o.ids, o.get(of))), tol)


@pytest.mark.parametrize("axis", [1, 0])
def test_mdbc_in_slab_mode_world_of_one(oracle_lib, axis):
    """SimpleMDBC through the slab-mode path (global ghost-node table, solves all-reduced, extrapolation by
    particle ID) against the per-particle single-GPU path and the oracle (C5, moving state)"""
    case = util.perturb(util.case_c5("float64"))
    p = util.params_of(case)
    assert p.mdbc == 1
    ref = Simulation(p)
    ref.upload(case.particles)
    r0 = ref.step(60, reset_delta_x=True)
    a = ref.download(order="id")
    ref.close()
    sim = Simulation(p)
    dec = slab.SlabDecomposition(sim, case.particles, p.H_inv, 0, 1, axis=axis).setup()
    gp, gid = dec.ghost_node_table()
    assert gp.shape[0] > 500 and np.all(np.diff(gid) > 0)
    r1 = sim.step(60, reset_delta_x=True)
    b = dec.gather(order="id", fields=("Position", "Velocity", "Density", "Pressure", "ID"))
    sim.close()
    o = oracle_lib.Oracle(p, case.particles, nthreads=4)
    o.step(60, True)
    assert r0["n_rebuilds"] == r1["n_rebuilds"] == o.report()["n_rebuilds"]
    bnd = np.asarray(case.particles.Type)[np.argsort(case.particles.ID, kind="stable")] != 1
    assert np.any(np.abs(b["Density"][bnd] - 1000.0) > 1e-3)       # the correction did something
    for k, of, tol in (("Position", "pos", 1e-13), ("Velocity", "vel", 1e-8), ("Density", "rho", 1e-11)):
        if axis == 1:   # same cell order as the plain run: same traversal, same sums
            assert np.array_equal(a[k], b[k]), k
        util.check(util.relerr(b[k], a[k]), tol)
        util.check(util.relerr(b[k], util.by_id(o.ids, o.get(of))), tol * 10)


def test_mdbc_in_slab_mode_needs_the_node_table():
    from sphexample_b200.simulation import SphError
    case = util.case_c5("float64")
    p = util.params_of(case)
    sim = Simulation(p)
    with pytest.raises(SphError):
        sim.set_ghost_nodes(np.zeros((1, 2)), np.ones(1, np.int64))    # before comm_init
    dec = slab.SlabDecomposition(sim, case.particles, p.H_inv, 0, 1, axis=1).join()
    with pytest.raises(SphError):
        sim.set_ghost_nodes(np.zeros((2, 2)), np.array([5, 5], np.int64))   # IDs not strictly ascending
    sim.upload(case.particles)
    with pytest.raises(SphError):
        sim.step(1, reset_delta_x=True)                                 # no table: refuse, do not skip S6
    sim.close()


def test_slab_mode_rejects_unsupported_setups():
    from sphexample_b200.simulation import SphError, comm_unique_id
    case = util.case_3d_small("float32")
    p = util.params_of(case)
    sim = Simulation(p)
    with pytest.raises(SphError):
        sim.comm_init(comm_unique_id(), 0, 1, 3)        # no such axis
    with pytest.raises(SphError):
        sim.set_slab(0, 10)                             # before comm_init
    sim.close()


def test_two_gpu_slab_parity(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run scripts/slab_parity.py under gpurun --gpus 2)")
    out = tmp_path / "slab.jsonl"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "slab_parity.py"), str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    rows = [json.loads(l) for l in open(out)]
    assert rows and all(r["ok"] for r in rows), rows
