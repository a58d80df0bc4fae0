#!/usr/bin/env python
"""bench.py — Mparticle-updates/s of the SPH inner loop on the 3D dam break (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU)

A "step" is one full symplectic SimulationLoop iteration (src/SPHCellList.jl:742-802) of the whole
domain: Δt/Δx reductions, (amortised) neighbour rebuild, two fused interaction passes with the
half / full updates.  Workload: config C3 of BASELINE.md — the 3D dam break regenerated on a
lattice with ~1 M particles per GPU (weak scaling: N GPUs -> ~N M particles, y-slab decomposition),
fp32 storage/compute, constants of example/Dambreak3d.jl.  Inputs are synthetic (deterministic
lattice, no RNG).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mparticle-updates/s (3D dam-break)"
UNIT = "Mparticle-updates/s"
PER_GPU_PARTICLES = 1_000_000
SLAB_AXIS = 1   # y: the dam break is (nearly) uniform along y, so y-slabs stay balanced through the run


def build_case(n_target, float_type="float32"):
    from sphexample_b200 import cases
    dp = cases.dp_for_count_3d(int(n_target))
    return cases.case_dam_break_3d(dp, float_type), dp


def params_of(case):
    from sphexample_b200 import make_params
    return make_params(case.meta, case.consts, case.kernel, case.viscosity, case.diffusion)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_rate(case, threads, steps, warmup=1):
    """the oracle (C++ restatement of the reference algorithm) on the host cores -> Mpu/s"""
    from oracle import oracle as orc
    p = params_of(case)
    o = orc.Oracle(p, case.particles, nthreads=threads)
    o.step(max(1, warmup), True)           # includes the first rebuild
    t0 = time.perf_counter()
    o.step(steps, False)
    dt = time.perf_counter() - t0
    n = len(case.particles)
    o.close()
    return n * steps / dt / 1e6, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia is not in this image)
    on all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    budget_s = float(os.environ.get("SPHB200_REF_BUDGET_S", "100"))
    # calibrate on a small lattice, then size the sample so that K + W steps fit the budget
    cal_case, _ = build_case(60_000, "float64")
    rate, _ = cpu_port_rate(cal_case, threads, 2)
    total_steps = args.steps + args.warmup
    n_full = PER_GPU_PARTICLES * max(1, args.gpus)
    n_sample = int(min(n_full, max(50_000, rate * 1e6 * budget_s / total_steps)))
    case, dp = build_case(n_sample, "float64")
    n = len(case.particles)
    o = orc.Oracle(params_of(case), case.particles, nthreads=threads)
    o.step(max(1, args.warmup), True)
    t0 = time.perf_counter()
    o.step(args.steps, False)
    el = time.perf_counter() - t0
    value = n * args.steps / el / 1e6
    sample = f"{n} particles (dp={dp}) of the {n_full}-particle workload, {args.steps} steps, fp64"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n_full, args.gpus, note="reference arm: C++ restatement of the reference's "
                                      "multi-threaded CPU algorithm (Julia unavailable), bounded sample, from rest "
                                      "(CPU cost per step does not depend on the flow state)", prep_time=args.prep_time),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_particles, gpus, note=None, prep_time=0.0):
    cfg = {"workload": f"C3 3D dam-break lattice (example/Dambreak3d.jl constants, h=sqrt(3)dp, Wendland C2, "
                       f"artificial viscosity + linear density diffusion), ~{PER_GPU_PARTICLES} particles per GPU, "
                       f"state {prep_time:g} s after release (developed flow; 0 = at rest)",
           "particles": int(n_particles), "precision": "fp32 storage+compute",
           "parallelism": "single GPU" if gpus == 1 else f"y-slab decomposition over {gpus} GPUs, NCCL halo exchange",
           "l2_policy": "L2 flushed between timed steps (256 MiB memset outside the per-step CUDA-event pairs); "
                        "value = particles*K / sum of per-step device times",
           "rebuild": "one forced neighbour rebuild at the start of the timed region + displacement-triggered ones "
                      "(reference cadence: forced rebuild per output interval of ~200 steps)"}
    if note:
        cfg["note"] = note
    return cfg


def max_over_ranks_host(torch, world, v):
    if world == 1:
        return v
    t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def run_ours(args, rank, world, local_rank):
    import torch
    from sphexample_b200.simulation import Simulation
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libsphb200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = torch.distributed if world > 1 else None
    n_total = PER_GPU_PARTICLES * world
    if args.particles:
        n_total = int(args.particles)
    case, dp = build_case(n_total, "float32")
    parts = case.particles
    n = len(parts)
    p = params_of(case)
    sim = Simulation(p, device=local_rank)
    stream = torch.cuda.Stream()              # a real stream (the legacy default stream cannot be graph-captured)
    torch.cuda.set_stream(stream)
    sim.set_stream(stream.cuda_stream)
    dec = None
    if world > 1 or os.environ.get("SPHB200_BENCH_FORCE_SLAB"):   # (world of one: the slab code path without neighbours)
        from sphexample_b200 import slab
        dec = slab.SlabDecomposition(sim, parts, p.H_inv, rank, world, axis=SLAB_AXIS)
        dec.join()
        mine = dec.mine
    else:
        mine = slice(None)
    # ---- workload state: the dam break t_prep seconds after release (developed flow) -------------
    # generated by running the case itself from rest (untimed, part of building the synthetic input);
    # the resulting per-rank particle table, in pinned HOST memory, is the benchmark's input
    sim.upload_arrays(*(np.ascontiguousarray(getattr(parts, k)[mine]) for k in ("Position", "Velocity", "Density")),
                      np.ascontiguousarray(parts.Type[mine], np.uint8), ids=np.ascontiguousarray(parts.ID[mine], np.int64))
    prep = {"t_prep": args.prep_time, "steps": 0}
    if args.prep_time > 0:
        r = sim.SimulationLoop(args.prep_time)
        prep["steps"] = int(r["iteration"])
        sim.set_time(0.0, 0)
    st0 = sim.download(fields=("Position", "Velocity", "Density", "Type", "ID"))
    prep["vmax"] = float(max_over_ranks_host(torch, world, np.sqrt((st0["Velocity"].astype(np.float64) ** 2).sum(1)).max()))
    host = {k: torch.from_numpy(np.ascontiguousarray(st0[k])).pin_memory().numpy() for k in ("Position", "Velocity", "Density")}
    types = np.ascontiguousarray(st0["Type"], np.uint8)
    ids = np.ascontiguousarray(st0["ID"], np.int64)
    n_local = int(types.shape[0])

    def upload():
        sim.upload_arrays(host["Position"], host["Velocity"], host["Density"], types, ids=ids)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def sum_over_ranks(v):
        if world == 1:
            return int(v)
        t = torch.tensor([int(v)], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    upload()
    # ---- warm-up -------------------------------------------------------------------------------
    warm = max(args.warmup, 3)
    sim.step(warm, reset_delta_x=True)
    barrier()
    # ---- timed region: exactly K steps, device events, max over ranks --------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # L2 policy: the per-GPU state (~84 B/particle) fits the 126 MB L2, so L2 is FLUSHED between
    # timed steps (a 256 MiB device memset outside the per-step event pairs); `value` is the sum
    # of the K per-step device times.  The un-flushed back-to-back figure is reported beside it.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    l0 = sim.launch_count
    rb0, lb0 = int(sim.report()["n_rebuilds"]), int(sim.stat("list_builds"))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    rep = None
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record(stream)
        rep = sim.step(1, reset_delta_x=(k == 0))
        ev[k][1].record(stream)
    barrier()
    t1 = time.time()
    ms = float(sum(a.elapsed_time(b) for a, b in ev))
    launches = sim.launch_count - l0
    rebuilds_timed = int(rep["n_rebuilds"]) - rb0
    list_builds_timed = int(sim.stat("list_builds")) - lb0
    # back-to-back (warm L2, one host sync per 64 steps on one GPU): what a production run sees
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    sim.step(args.steps, reset_delta_x=True)
    e1.record(stream)
    barrier()
    ms_b2b = e0.elapsed_time(e1)
    ms, ms_b2b = max_over_ranks([ms, ms_b2b])
    launches = sum_over_ranks(launches)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = n * args.steps / (ms * 1e-3) / 1e6

    # ---- per-kernel times for the roofline (device events around the stages of extra steps) ----
    stage = np.zeros(5)
    reps = 5
    for _ in range(reps):
        flush.zero_()
        stage += np.array(sim.stage_times())
    stage = np.array(max_over_ranks(stage / reps))
    n_local_max = int(max_over_ranks([sim.report()["n_particles"]])[0])
    D, sz = 3, 4
    bytes_pass0 = n_local_max * (4 * D + 4) * sz          # read x,v,rho,P ; write x_h,v_h,rho_h,P_h
    bytes_pass1 = n_local_max * (7 * D + 5) * sz          # read half state + own state n + rho_n ; write x,v,rho,P,a
    peak, how = measured_peaks()
    t_avg = 0.5 * (stage[2] + stage[3]) * 1e-3
    achieved = 0.5 * (bytes_pass0 + bytes_pass1) / t_avg / 1e9
    prof = os.path.join(ROOT, "profiles", "interact_traffic.json")
    traffic = json.load(open(prof)).get("dram_bytes_per_launch") if os.path.exists(prof) else None
    roof = {"bound": "hbm", "kernel": "k_interact_list<float,3,PASS> (+ k_list_build / k_interact when a pass cannot use the lists); avg of the two passes of a step, slowest rank",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": how, "algorithmic_bytes_per_launch": 0.5 * (bytes_pass0 + bytes_pass1),
            "avg_launch_ms": 0.5 * (stage[2] + stage[3]),
            "note": "not an HBM-bound kernel (~18 kflop per particle per pass, SURVEY 8d): bound by shared-memory gather "
                    "bandwidth and fp32 issue slots (profiles/); the HBM fraction is reported because the metric asks for it",
            "stage_ms": {"reduce_control": stage[0], "rebuild_predicated": stage[1], "pass0_fused": stage[2],
                         "pass1_fused": stage[3], ("metadata" if world == 1 else "halo_exchanges"): stage[4]}}

    # ---- end to end through the reference-facing call with HOST buffers -------------------------
    cap = n_local + n_local // 8 + 1024            # owned counts drift a little with migration
    out = {k: torch.empty((cap,) + v.shape[1:], dtype=torch.float32).pin_memory().numpy() for k, v in
           (("Position", host["Position"]), ("Velocity", host["Velocity"]), ("Density", host["Density"]),
            ("Pressure", host["Density"]))}
    iters = 2
    barrier()
    e0.record(stream)
    for _ in range(iters):
        upload()
        sim.step(args.steps, reset_delta_x=True)
        sim.download_into(out["Position"][:sim.num_particles], out["Velocity"][:sim.num_particles],
                          out["Density"][:sim.num_particles], out["Pressure"][:sim.num_particles])
    e1.record(stream)
    barrier()
    ems = max_over_ranks([e0.elapsed_time(e1)])[0]
    h2d = sum_over_ranks(sum(host[k].nbytes for k in host) + types.nbytes + ids.nbytes)
    d2h = sum_over_ranks(sum(v[:sim.num_particles].nbytes for v in out.values()))
    e2e = {"value": n * args.steps * iters / (ems * 1e-3) / 1e6, "unit": UNIT,
           "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
           "definition": f"one SimulationLoop-style call per output interval: upload host particle table "
                         f"({h2d} B, pinned) -> {args.steps} steps -> download x,v,rho,P ({d2h} B); "
                         f"bytes amortised per step, all ranks"}
    # worst case for context: a host round trip around EVERY step
    k2 = 5
    barrier()
    e0.record(stream)
    for _ in range(k2):
        upload()
        sim.step(1, reset_delta_x=True)
        sim.download_into(out["Position"][:sim.num_particles], out["Velocity"][:sim.num_particles],
                          out["Density"][:sim.num_particles], out["Pressure"][:sim.num_particles])
    e1.record(stream)
    barrier()
    e2e["roundtrip_every_step_value"] = n * k2 / (max_over_ranks([e0.elapsed_time(e1)])[0] * 1e-3) / 1e6

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(n, world, prep_time=args.prep_time), "clocks": clocks,
                "gpu_launches": int(launches), "roofline": roof, "e2e": e2e,
                "rebuilds_in_timed_region": rebuilds_timed, "list_builds_in_timed_region": list_builds_timed, "dp": dp,
                "state": prep,
                "back_to_back": {"value": n * args.steps / (ms_b2b * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_b2b / args.steps,
                                 "note": "same K steps enqueued back to back, warm L2"}}
        if world > 1:
            line["slab"] = {"axis": "xyz"[SLAB_AXIS], "edges": [int(e) for e in dec.edges],
                            "owned_max": n_local_max, "owned_mean": n / world}
        # ---- CPU baseline (oracle port) on a bounded sample of the same workload, N = 1 only ----
        if world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline_sample(n_total)
            except Exception as ex:   # the checker failing must not void the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    sim.close()


def cpu_baseline_sample(n_total, budget_s=None):
    """the oracle port on all host threads: the SAME workload, a bounded number of steps (~15 s)"""
    from oracle import oracle as orc
    orc.build()
    threads = orc.max_threads()
    budget_s = float(os.environ.get("SPHB200_CPU_BUDGET_S", "15")) if budget_s is None else budget_s
    case, dp = build_case(n_total, "float64")
    o = orc.Oracle(params_of(case), case.particles, nthreads=threads)
    t0 = time.perf_counter()
    o.step(1, True)                        # warm-up step incl. the first rebuild
    t_first = time.perf_counter() - t0
    steps = int(max(2, min(50, budget_s / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    o.step(steps, False)
    secs = time.perf_counter() - t0
    npart = len(case.particles)
    o.close()
    return {"value": npart * steps / secs / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{steps} steps (after 1 warm-up step incl. rebuild) of the same {npart}-particle workload "
                      f"(dp={dp}), fp64, {secs:.1f} s on {threads} host threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=0, help="override the total particle count")
    ap.add_argument("--prep-time", type=float, default=0.15,
                    help="simulated seconds the dam break runs (untimed) before the benchmark state is taken; 0 = from rest")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # watchdog: a wedged run (a peer rank died, a collective never completes) must end, not hang its caller
    limit = float(os.environ.get("SPHB200_BENCH_TIMEOUT_S", "1500"))

    def _expired():
        sys.stderr.write(f"bench.py: rank {rank} exceeded {limit:.0f} s, giving up\n")
        sys.stderr.flush()
        os._exit(3)
    wd = threading.Timer(limit, _expired)
    wd.daemon = True
    wd.start()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
