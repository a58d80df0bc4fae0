"""Multi-GPU parity: the slab-decomposed run (one rank per GPU, NCCL halo exchange inside the
library) against the single-GPU run of the same library and against the CPU oracle.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 scripts/slab_parity.py [out.jsonl]

Every rank builds the same deterministic case; rank 0 also runs the references and prints one JSON
line per (case, precision, axis) with the measured relative errors (by particle ID)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from sphexample_b200 import slab  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out_path = sys.argv[1] if len(sys.argv) > 1 else None
lines = []


def run_case(name, make, steps, axis, tol_single, tol_oracle, vel_scale=0.5):
    case = util.perturb(make(), vel_scale=vel_scale)
    p = util.params_of(case)
    try:
        slab.plan_edges(slab.cell_coord(case.particles.Position[:, axis], p.H_inv), world)
    except ValueError as ex:      # too few cell layers along this axis for `world` slabs: same verdict on every rank
        if rank == 0:
            print(json.dumps({"case": name, "axis": axis, "world": world, "skipped": str(ex)}), flush=True)
        return
    sim = Simulation(p, device=local)
    dec = slab.SlabDecomposition(sim, case.particles, p.H_inv, rank, world, axis=axis).setup()
    rep = sim.step(steps, reset_delta_x=True)
    st = dec.gather(order="id", fields=("Position", "Velocity", "Density", "Pressure", "ID"))
    counts = [None] * world
    dist.all_gather_object(counts, (int(rep["n_particles"]), int(rep["n_halo"]), int(rep["n_rebuilds"])))
    sim.close()
    if rank == 0:
        ref = Simulation(p, device=local)
        ref.upload(case.particles)
        rref = ref.step(steps, reset_delta_x=True)
        s1 = ref.download(order="id")
        ref.close()
        from oracle import oracle as orc
        o = orc.Oracle(p, case.particles, nthreads=8)
        o.step(steps, True)
        ids = o.ids
        rec = {"case": name, "float": case.meta.FloatType, "axis": dec.axis, "world": world, "steps": steps,
               "edges": [int(e) for e in dec.edges], "owned_halo_rebuilds": counts,
               "n_total": int(sum(c[0] for c in counts)), "n_expected": len(case.particles),
               "time_slab": rep["total_time"], "time_single": rref["total_time"], "ids_equal": bool(np.array_equal(st["ID"], s1["ID"]))}
        for f, of in (("Position", "pos"), ("Velocity", "vel"), ("Density", "rho")):
            rec[f"err_{f}_vs_single"] = util.relerr(st[f], s1[f])
            rec[f"err_{f}_vs_oracle"] = util.relerr(st[f], util.by_id(ids, o.get(of)))
        rec["ok"] = bool(rec["n_total"] == rec["n_expected"] and rec["ids_equal"] and
                         max(rec[f"err_{f}_vs_single"] for f in ("Position", "Velocity", "Density")) < tol_single and
                         max(rec[f"err_{f}_vs_oracle"] for f in ("Position", "Velocity", "Density")) < tol_oracle)
        rec["tol_single"], rec["tol_oracle"] = tol_single, tol_oracle
        lines.append(rec)
        print(json.dumps(rec), flush=True)
    dist.barrier()


if os.environ.get("SLAB_PARITY_QUICK"):   # a short smoke of the exchange / migration / pause paths
    run_case("dam_break_3d_dp0.02", lambda: util.case_3d_small("float32"), 30, 1, 5e-3, 5e-3)
    run_case("dam_break_3d_dp0.02_fast", lambda: util.case_3d_small("float64"), 60, 1, 1e-7, 1e-6, vel_scale=3.0)
    if rank == 0:
        print("SLAB PARITY", "OK" if all(r["ok"] for r in lines) else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
def mdbc_cases():
    # C5 (StillWedge, SimpleMDBC): ghost nodes lie up to two cells from their particles, across the faces
    run_case("still_wedge_mdbc_2d_z", lambda: util.case_c5("float64"), 80, 1, 1e-9, 1e-7)
    run_case("still_wedge_mdbc_2d_x", lambda: util.case_c5("float64"), 80, 0, 1e-9, 1e-7)
    run_case("still_wedge_mdbc_2d_x_fast", lambda: util.case_c5("float64"), 150, 0, 1e-8, 1e-6, vel_scale=3.0)


if os.environ.get("SLAB_PARITY_ONLY") == "mdbc":
    mdbc_cases()
    if rank == 0:
        if out_path:
            os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
            with open(out_path, "w") as fh:
                for r in lines:
                    fh.write(json.dumps(r) + "\n")
        print("SLAB PARITY", "OK" if lines and all(r["ok"] for r in lines) else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
axes3 = (0, 1, 2)     # x-slabs (rows along y), y-slabs, z-slabs
for ax in axes3:
    run_case("dam_break_3d_dp0.02", lambda: util.case_3d_small("float64"), 60, ax, 1e-9, 1e-8)
    run_case("dam_break_3d_dp0.02", lambda: util.case_3d_small("float32"), 60, ax, 5e-3, 5e-3)
# fast flow: many rebuilds and migrations across the slab faces
run_case("dam_break_3d_dp0.02_fast", lambda: util.case_3d_small("float64"), 150, 1, 1e-7, 1e-6, vel_scale=3.0)
run_case("dam_break_3d_dp0.02_fast_x", lambda: util.case_3d_small("float64"), 150, 0, 1e-7, 1e-6, vel_scale=3.0)
run_case("dam_break_2d_dp0.02", lambda: util.case_c1("float64"), 100, 1, 1e-9, 1e-8)
run_case("dam_break_2d_dp0.02_x", lambda: util.case_c1("float64"), 100, 0, 1e-9, 1e-8)
mdbc_cases()
if rank == 0:
    if out_path:
        os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
        with open(out_path, "w") as fh:
            for r in lines:
                fh.write(json.dumps(r) + "\n")
    print("SLAB PARITY", "OK" if all(r["ok"] for r in lines) else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
