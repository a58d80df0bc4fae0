"""the single-GPU configs C1, C2, C5 (2D fp64) and the shipped-size 3D case, each
with GPU Mpu/s (back-to-back steps, CUDA events), parity vs the oracle after the same steps, and the
oracle's own rate at 1 thread and at all host threads.
  python scripts/config_table.py > gpurun_out/<tag>/configs.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from sphexample_b200 import cases  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

orc.build()
THREADS = orc.max_threads()


def gpu_rate(case, steps, warm, **opts):
    p = util.params_of(case)
    sim = Simulation(p)
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.upload(case.particles)
    sim.step(warm, reset_delta_x=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.Stream()          # a real stream: the legacy default stream cannot be graph-captured
    sim.set_stream(stream.cuda_stream)
    sim.step(2)                           # (captures the step graph; the oracle below takes these 2 steps too)
    torch.cuda.synchronize()
    e0.record(stream)
    sim.step(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = sim.download(order="id")
    rep = sim.report()
    builds = sim.stat("list_builds")
    sim.close()
    return len(case.particles) * steps / ms / 1e3, st, rep, builds


def cpu_rate(case, steps, warm, threads):
    p = util.params_of(case)
    o = orc.Oracle(p, case.particles, nthreads=threads)
    o.step(warm, True)
    t0 = time.perf_counter()
    o.step(steps, False)
    dt = time.perf_counter() - t0
    return len(case.particles) * steps / dt / 1e6, o


def row(name, case, steps, warm, cpu_steps, **opts):
    rate, st, rep, builds = gpu_rate(case, steps, warm, **opts)
    r1, _ = cpu_rate(case, cpu_steps, 1, 1)
    rN, _ = cpu_rate(case, cpu_steps, 1, THREADS)
    # parity after warm + steps steps (same cadence: one forced rebuild at the start)
    o = orc.Oracle(util.params_of(case), case.particles, nthreads=THREADS)
    o.step(warm, True)
    o.step(steps + 2, False)
    ids = o.ids
    out = {"config": name, "n": len(case.particles), "float": case.meta.FloatType, "opts": opts, "steps": steps,
           "gpu_Mpu_s": round(rate, 2), "cpu_1thread_Mpu_s": round(r1, 4), "cpu_all_Mpu_s": round(rN, 4), "cpu_threads": THREADS,
           "rebuilds": int(rep["n_rebuilds"]), "list_builds": builds,
           "err_vel": util.relerr(st["Velocity"], util.by_id(ids, o.get("vel"))),
           "err_rho": util.relerr(st["Density"], util.by_id(ids, o.get("rho"))),
           "err_pos": util.relerr(st["Position"], util.by_id(ids, o.get("pos")))}
    print(json.dumps(out), flush=True)


P = lambda c: util.perturb(c, vel_scale=1.0)   # a moving state: from rest the velocities are ~0 and relative errors meaningless
row("C1 2D dam break shipped (6 881), fp64", P(util.case_c1("float64")), 400, 20, 40)
row("C1 full step sequence, whole-list lanes (round-2 start)", P(util.case_c1("float64")), 400, 20, 40, lean=0, split=0, list_local=1)
row("C2 2D dam break dp=0.0058 (59 909), fp64", P(cases.case_dam_break_2d(0.0058, "float64")), 400, 20, 10)
row("C5 StillWedge mDBC (3 027), fp64", util.case_c5("float64"), 400, 20, 40)
row("3D dam break, the shipped 171 496-particle files, fp32", P(util.case_3d_shipped("float32")), 200, 20, 5)
