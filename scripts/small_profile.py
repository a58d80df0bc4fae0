"""Where a step of the SMALL configs (C1, C2, C5: launch-latency bound) spends its time: rate over 400
steps (step graph), per-stage device times (median of 30 sphb200_stage_times calls), launches per step.
  python scripts/small_profile.py [c1|c2|c5 ...] > gpurun_out/<tag>/small_profile.jsonl
Under ncu (launch list) use SMALL_PROFILE_NCU=1: 40 plain-launch steps only."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from sphexample_b200 import cases  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

P = lambda c: util.perturb(c, vel_scale=1.0)
CASES = {"c1": lambda: P(util.case_c1("float64")), "c2": lambda: P(cases.case_dam_break_2d(0.0058, "float64")),
         "c5": lambda: util.case_c5("float64")}
NCU = bool(os.environ.get("SMALL_PROFILE_NCU"))
opts = {k[4:].lower(): float(v) for k, v in os.environ.items() if k.startswith("OPT_")}

for name in (sys.argv[1:] or ["c1", "c2", "c5"]):
    case = CASES[name]()
    sim = Simulation(util.params_of(case))
    for k, v in opts.items():
        sim.set_option(k, v)
    if NCU:
        sim.set_option("graph", 0)
    sim.upload(case.particles)
    stream = torch.cuda.Stream()
    sim.set_stream(stream.cuda_stream)
    sim.step(20, reset_delta_x=True)
    if NCU:
        sim.step(40)
        torch.cuda.synchronize()
        sim.close()
        continue
    sim.step(2)
    torch.cuda.synchronize()
    l0 = sim.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(400)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rep = sim.report()
    st = [sim.stage_times() for _ in range(30)]
    med = {k: round(float(np.median([s[k] for s in st])) * 1e3, 2) for k in st[0]}
    print(json.dumps({"config": name, "n": len(case.particles), "opts": opts, "Mpu_s": round(len(case.particles) * 400 / ms / 1e3, 2),
                      "us_per_step": round(ms / 400 * 1e3, 2), "launches_per_step": (sim.launch_count - l0) / 400 if False else None,
                      "rebuilds": int(rep["n_rebuilds"]), "list_builds": sim.stat("list_builds"),
                      "stage_us_median": med, "stage_sum_us": round(sum(med.values()), 1)}), flush=True)
    sim.close()
