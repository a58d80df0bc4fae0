"""The lean step sequence (steps that will not rebuild replay a graph without the UpdateNeighbors! chain and
the reductions sweep; k_step_control pauses a step that turns out to need the chain): same kernels on the
same data in the same order as the full sequence, so the results must be bitwise identical."""
import numpy as np
import pytest
import torch

import util
from sphexample_b200 import cases
from sphexample_b200.simulation import Simulation

pytestmark = pytest.mark.gpu


def _run(case, steps, chunks, **opts):
    sim = Simulation(util.params_of(case))
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.upload(case.particles)
    stream = torch.cuda.Stream()          # a capturable stream: the step graphs are used
    sim.set_stream(stream.cuda_stream)
    rep = None
    for k, n in enumerate(chunks):
        rep = sim.step(n, reset_delta_x=(k == 0))
    assert rep["iteration"] == steps
    st = sim.download(order="id")
    stats = {k: sim.stat(k) for k in ("lean_steps", "lean_pauses", "list_builds")}
    sim.close()
    return st, rep, stats


@pytest.mark.parametrize("name", ["c1_fast", "c5", "3d_f32"])
def test_lean_sequence_is_bitwise_the_full_sequence(name):
    mk = {"c1_fast": lambda: util.perturb(util.case_c1("float64"), vel_scale=3.0),
          "c5": lambda: util.perturb(util.case_c5("float64")),
          "3d_f32": lambda: util.perturb(util.case_3d_small("float32"), vel_scale=2.0)}[name]
    a, ra, sa = _run(mk(), 150, [150], lean=0)
    b, rb, sb = _run(mk(), 150, [150], lean=1)
    c, rc, sc = _run(mk(), 150, [1, 7, 1, 41, 100], lean=1, batch=7)     # odd call / batch boundaries
    assert sa["lean_steps"] == 0 and sb["lean_steps"] > 100 and sc["lean_steps"] > 100
    assert ra["n_rebuilds"] == rb["n_rebuilds"] == rc["n_rebuilds"] >= 2
    assert ra["total_time"] == rb["total_time"] == rc["total_time"]
    for k in ("Position", "Velocity", "Density", "Pressure", "Acceleration"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
    assert sa["list_builds"] == sb["list_builds"] == sc["list_builds"]


def test_a_mispredicted_lean_step_pauses_and_is_finished_with_the_rebuild():
    """batches far longer than the rebuild interval predicted from a slow start: the flow accelerates
    (vel_scale grows the displacement per step), so some lean step must find delta_x used up"""
    case = util.perturb(util.case_c1("float64"), vel_scale=0.05)
    case.particles.Velocity[case.particles.Type == 1] *= 1.0
    a, ra, sa = _run(case, 600, [600], lean=0)
    case = util.perturb(util.case_c1("float64"), vel_scale=0.05)
    b, rb, sb = _run(case, 600, [600], lean=1, batch=64)
    assert ra["n_rebuilds"] == rb["n_rebuilds"]
    for k in ("Position", "Velocity", "Density"):
        assert np.array_equal(a[k], b[k]), k
    assert sb["lean_steps"] > 400


def test_a_list_build_that_fails_inside_a_lean_step_pauses_it_before_the_passes():
    """the lean sequence has no cull kernels standing by: when the step's own list build overflows, the list
    kernel stops the step (ctl->paused = 2) and the host runs the two passes with the full sequence — the cull
    kernels — after which the lists stay off until the next cell rebuild (test hook: the overflow is injected)"""
    mk = lambda: util.perturb(util.case_c1("float64"), vel_scale=1.0)
    a, ra, sa = _run(mk(), 60, [60], lean=0)
    sim = Simulation(util.params_of(mk()))
    for k, v in (("lean", 1), ("graph", 0), ("test_fail_list_build_at", 12)):
        sim.set_option(k, v)
    sim.upload(mk().particles)
    rep = sim.step(60, reset_delta_x=True)
    b = sim.download(order="id")
    assert sim.stat("lean_pauses") >= 1 and sim.stat("list_fail_reason") == 2 and sim.stat("lean_steps") >= 12
    sim.close()
    assert rep["iteration"] == 60 and rep["n_rebuilds"] == ra["n_rebuilds"] and rep["total_time"] == pytest.approx(ra["total_time"], rel=1e-12)
    for k in ("Position", "Velocity", "Density"):
        util.check(util.relerr(b[k], a[k]), 1e-10)      # cull and list kernels sum the same pairs in another order


def test_simulation_loop_target_time_with_the_lean_sequence():
    case = util.perturb(util.case_c1("float64"), vel_scale=1.0)
    out = []
    for lean in (0, 1):
        sim = Simulation(util.params_of(case))
        sim.set_option("lean", lean)
        sim.upload(case.particles)
        stream = torch.cuda.Stream()
        sim.set_stream(stream.cuda_stream)
        r1 = sim.SimulationLoop(0.002)
        r2 = sim.SimulationLoop(0.004)
        out.append((r1["iteration"], r2["iteration"], r2["total_time"], sim.download(order="id")))
        assert r2["total_time"] > 0.004 and r1["total_time"] > 0.002
        sim.close()
    assert out[0][:3] == out[1][:3]
    for k in ("Position", "Velocity", "Density"):
        assert np.array_equal(out[0][3][k], out[1][3][k]), k
