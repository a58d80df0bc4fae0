"""Option sweep on the bench workload in DEVELOPED flow (the state bench.py times): the dam break is
run once to t_prep, the state is kept on the host, and every option set is timed from that state.
  python scripts/tune.py [n_particles] [t_prep]     env SPH_SWEEP=';'-separated 'k=v,k=v' sets
One JSON line per option set: ms/step (back-to-back, CUDA events), Mpu/s, list builds, cell rebuilds."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
t_prep = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
steps = int(os.environ.get("SPH_STEPS", "120"))
dflt = ("lists=1;lists=1,list_reorder=0;lists=1,skin=0.07;lists=1,skin=0.15;lists=0")
sets = os.environ.get("SPH_SWEEP", dflt).split(";")
case, dp = bench.build_case(n, "float32")
p = bench.params_of(case)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)

sim = Simulation(p)
sim.set_stream(stream.cuda_stream)
sim.upload(case.particles)
if t_prep > 0:
    sim.SimulationLoop(t_prep)
st = sim.download(fields=("Position", "Velocity", "Density", "Type", "ID"))
sim.close()
vmax = float(np.sqrt((st["Velocity"].astype(np.float64) ** 2).sum(1)).max())
types = np.ascontiguousarray(st["Type"], np.uint8)
ids = np.ascontiguousarray(st["ID"], np.int64)
for s in sets:
    opts = dict(kv.split("=") for kv in s.split(",") if kv)
    sim = Simulation(p)
    for k, v in opts.items():
        sim.set_option(k, float(v))
    sim.set_stream(stream.cuda_stream)
    sim.upload_arrays(st["Position"], st["Velocity"], st["Density"], types, ids=ids)
    sim.step(12, reset_delta_x=True)
    r0, b0 = sim.report(), sim.stat("list_builds")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    sim.step(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    r1 = sim.report()
    print(json.dumps({"opts": opts, "n": len(ids), "t_prep": t_prep, "vmax": round(vmax, 3), "ms_per_step": round(ms, 4),
                      "Mpu_s": round(len(ids) / ms / 1e3, 1), "list_builds": sim.stat("list_builds") - b0,
                      "cell_rebuilds": r1["n_rebuilds"] - r0["n_rebuilds"], "steps": steps, "list_off": sim.stat("list_off"),
                      "list_wavefronts": round(sim.stat("list_wavefronts"), 3), "list_entries": round(sim.stat("list_entries"), 1)}), flush=True)
    sim.close()
