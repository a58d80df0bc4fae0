"""Where the end-to-end leg of bench.py spends its time beyond the steps: raw pinned-memory PCIe rates on this
box, and the library's upload / download calls timed on their own (C3-sized table, fp32)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

out = {}
for mb in (4, 12, 37):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        out[f"{name}_{mb}MB_GBs"] = round(10 * (mb << 20) / (time.perf_counter() - t0) / 1e9, 2)
case, dp = bench.build_case(1_000_000, "float32")
P = case.particles
sim = Simulation(bench.params_of(case))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sim.set_stream(stream.cuda_stream)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
host = {k: pin(getattr(P, k)) for k in ("Position", "Velocity", "Density")}
types, ids = pin(P.Type.astype(np.uint8)), pin(P.ID.astype(np.int64))
n = len(P)
o = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory().numpy() for k, v in
     (("Position", host["Position"]), ("Velocity", host["Velocity"]), ("Density", host["Density"]), ("Pressure", host["Density"]))}
up = lambda: sim.upload_arrays(host["Position"], host["Velocity"], host["Density"], types, ids=ids)
dn = lambda: sim.download_into(o["Position"], o["Velocity"], o["Density"], o["Pressure"])
up(); sim.step(3, reset_delta_x=True); dn()
for name, fn in (("upload_ms", up), ("step1_after_upload_ms", lambda: sim.step(1, reset_delta_x=True)), ("step1_ms", lambda: sim.step(1)),
                 ("step18_ms", lambda: sim.step(18)), ("download_ms", dn)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    out[name] = round((time.perf_counter() - t0) * 1e3, 3)
out["upload_bytes"] = int(sum(v.nbytes for v in host.values()) + types.nbytes + ids.nbytes)
out["download_bytes"] = int(sum(v.nbytes for v in o.values()))
print(json.dumps(out))
