"""Cost of a FULL list build (after an upload / a cell rebuild): time of the first fused step after an upload
against a later no-rebuild step, for option sets given in SPH_SWEEP.   python scripts/time_build.py [n]"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
case, dp = bench.build_case(n, "float32")
P = case.particles
f = (P.Type == 1); x = P.Position.astype(np.float64)
P.Velocity[:, 0] = (1.5 * np.sin(3.0 * x[:, 2] + 1.0) * f).astype(np.float32)
P.Velocity[:, 2] = (-1.5 * np.cos(2.0 * x[:, 0]) * f).astype(np.float32)
p = bench.params_of(case)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for s in os.environ.get("SPH_SWEEP", "lists=1").split(";"):
    opts = dict(kv.split("=") for kv in s.split(",") if kv)
    sim = Simulation(p)
    for k, v in opts.items(): sim.set_option(k, float(v))
    sim.set_stream(stream.cuda_stream)
    first = []
    for rep in range(4):
        sim.upload(P)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize(); e0.record(stream); sim.step(1, reset_delta_x=True); e1.record(stream); sim.step(1); e2.record(stream); torch.cuda.synchronize()
        first.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    print(json.dumps({"opts": opts, "first_step_ms": round(min(a for a, b in first), 4), "second_step_ms": round(min(b for a, b in first), 4),
                      "list_wavefronts": round(sim.stat("list_wavefronts"), 3)}), flush=True)
    sim.close()
