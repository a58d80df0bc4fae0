"""A few steps of the bench workload (C3, ~1 M particles, fp32) for ncu / sweeps.
  python scripts/profile_step.py [n_particles] [steps]
env: SPH_PREP=<seconds>  run the dam break that long first (the developed flow bench.py times) and
     switch the profiler on only afterwards (use `ncu --profile-from-start off`);
     SPH_VEL=<m/s>       (without SPH_PREP) a smooth velocity field as a cheap developed-flow proxy;
     SPH_OPTS='k=v,k=v'  library options;
     SPH_RESET=1         re-arm the forced rebuild for the profiled steps."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ft = os.environ.get("SPH_FLOAT", "float32")
prep = float(os.environ.get("SPH_PREP", "0"))
case, dp = bench.build_case(n, ft)
sim = Simulation(bench.params_of(case))
for kv in os.environ.get("SPH_OPTS", "").split(","):
    if kv:
        k, v = kv.split("=")
        sim.set_option(k, float(v))
vel = float(os.environ.get("SPH_VEL", "1.5"))   # developed-flow proxy so that list builds occur
if prep <= 0 and vel > 0:
    P = case.particles
    f = (P.Type == 1)
    x = P.Position.astype(np.float64)
    P.Velocity[:, 0] = (vel * np.sin(3.0 * x[:, 2] + 1.0) * f).astype(P.Velocity.dtype)
    P.Velocity[:, 2] = (-vel * np.cos(2.0 * x[:, 0]) * f).astype(P.Velocity.dtype)
sim.upload(case.particles)
if prep > 0:
    sim.SimulationLoop(prep)
    sim.set_time(0.0, 0)
sim.step(3, reset_delta_x=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sim.step(steps, reset_delta_x=bool(os.environ.get("SPH_RESET")))   # SPH_RESET=1: the forced UpdateNeighbors! + full list build is in the profiled region
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", len(case.particles), dp, sim.report(), "list_wavefronts", sim.stat("list_wavefronts"), "list_entries", sim.stat("list_entries"))
