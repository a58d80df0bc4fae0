"""Option sweep of the interaction kernel on the bench workload: per-stage device times.
  python scripts/sweep.py [n_particles]   (env SPH_SWEEP = ';'-separated 'k=v,k=v' option sets)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
ft = os.environ.get("SPH_FLOAT", "float32")
dflt = "compact=1,tma=1;compact=0,tma=1;compact=1,tma=0;compact=1,tma=1,smem_kb=64;compact=1,tma=1,smem_kb=110;compact=1,tma=1,smem_kb=200"
sets = os.environ.get("SPH_SWEEP", dflt).split(";")
case, dp = bench.build_case(n, ft)
p = bench.params_of(case)
for s in sets:
    opts = dict(kv.split("=") for kv in s.split(",") if kv)
    sim = Simulation(p)
    for k, v in opts.items():
        sim.set_option(k, float(v))
    sim.upload(case.particles)
    sim.step(8, reset_delta_x=True)
    _t = [sim.stage_times() for _ in range(6)]
    _v = np.mean([list(t.values()) for t in _t], axis=0)
    st = [_v[0], _v[1] + _v[2] + _v[3], _v[4] + _v[5], _v[6] + _v[7], _v[8]]   # legacy 5 buckets: head, rebuild+motion, pass 1 (+lists), pass 2, metadata
    print(json.dumps({"opts": opts, "n": len(case.particles), "float": ft, "pass0_ms": st[2], "pass1_ms": st[3],
                      "reduce_ms": st[0], "rebuild_ms": st[1], "Mpu_s": len(case.particles) / (sum(st) * 1e-3) / 1e6}), flush=True)
    sim.close()
