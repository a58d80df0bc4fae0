"""List-kernel diagnostics on the bench workload: builds, overflow state, per-pass stage times.
  python scripts/list_diag.py [n] ;  env SPH_SWEEP as in sweep.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sphexample_b200.simulation import Simulation
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
sets = os.environ.get("SPH_SWEEP", "lists=1").split(";")
case, dp = bench.build_case(n, "float32")
vel = float(os.environ.get("SPH_VEL", "0"))
if vel > 0:   # developed-flow proxy: a smooth velocity field of amplitude `vel` m/s on the fluid
    P = case.particles
    f = (P.Type == 1)
    x = P.Position.astype(np.float64)
    P.Velocity[:, 0] = (vel * np.sin(3.0 * x[:, 2] + 1.0) * f).astype(np.float32)
    P.Velocity[:, 2] = (-vel * np.cos(2.0 * x[:, 0]) * f).astype(np.float32)
    P.Velocity[:, 1] = (0.3 * vel * np.sin(5.0 * x[:, 0] + 2.0 * x[:, 2]) * f).astype(np.float32)
p = bench.params_of(case)
for s in sets:
    opts = dict(kv.split("=") for kv in s.split(",") if kv)
    sim = Simulation(p)
    for k, v in opts.items():
        sim.set_option(k, float(v))
    sim.upload(case.particles)
    sim.step(8, reset_delta_x=True)
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _t = [sim.stage_times() for _ in range(6)]
    _v = np.mean([list(t.values()) for t in _t], axis=0)
    st = [_v[0], _v[1] + _v[2] + _v[3], _v[4] + _v[5], _v[6] + _v[7], _v[8]]   # legacy 5 buckets: head, rebuild+motion, pass 1 (+lists), pass 2, metadata
    b0 = sim.stat("list_builds"); r0 = sim.report()
    K = int(os.environ.get("SPH_STEPS", "60"))
    torch.cuda.synchronize(); e0.record(); sim.step(K); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    r1 = sim.report()
    print(json.dumps({"opts": opts, "n": len(case.particles), "vel": vel, "pass0_ms": round(st[2], 4), "pass1_ms": round(st[3], 4),
                      "ms_per_step_b2b": round(ms, 4), "Mpu_s": round(len(case.particles) / ms / 1e3, 1),
                      "list_builds_in_K": sim.stat("list_builds") - b0, "cell_rebuilds_in_K": r1["n_rebuilds"] - r0["n_rebuilds"],
                      "K": K, "list_off": sim.stat("list_off"), "fail_reason": sim.stat("list_fail_reason")}), flush=True)
    sim.close()
