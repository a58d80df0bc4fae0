"""A few steps of the bench workload (C3, ~1 M particles, fp32) for ncu / sweeps.
  python scripts/profile_step.py [n_particles] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ft = os.environ.get("SPH_FLOAT", "float32")
case, dp = bench.build_case(n, ft)
sim = Simulation(bench.params_of(case))
vel = float(os.environ.get("SPH_VEL", "1.5"))   # developed-flow proxy so that list builds occur
if vel > 0:
    import numpy as np
    P = case.particles
    f = (P.Type == 1)
    x = P.Position.astype(np.float64)
    P.Velocity[:, 0] = (vel * np.sin(3.0 * x[:, 2] + 1.0) * f).astype(P.Velocity.dtype)
    P.Velocity[:, 2] = (-vel * np.cos(2.0 * x[:, 0]) * f).astype(P.Velocity.dtype)
sim.upload(case.particles)
sim.step(3, reset_delta_x=True)
sim.step(steps)
print("ok", len(case.particles), dp, sim.report())
