#!/usr/bin/env python
"""bench.py — Mparticle-updates/s of the SPH inner loop on the 3D dam break (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU)

A "step" is one full symplectic SimulationLoop iteration (src/SPHCellList.jl:742-802) of the whole
domain: Δt/Δx reductions, (amortised) neighbour rebuild, two fused interaction passes with the
half / full updates.  Workload (BASELINE.md / SURVEY §8d), the 3D dam break regenerated on a lattice,
fp32 particle state, constants of example/Dambreak3d.jl, taken 0.15 s after release (developed flow):
  N = 1      C3: ~1.0 M particles on one GPU (the configuration the metric is quoted on);
  N = 2/4/8  C4 particle density: ~2.04 M particles per GPU in y-slabs — at N = 8 this IS C4
             (~16.3 M particles, dp ~ 0.0017).
Inputs are synthetic (deterministic lattice, no RNG).  One JSON line on stdout (rank 0).
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
C3_PARTICLES = 1_000_000
C4_PARTICLES = 16_320_000          # over 8 GPUs
SLAB_AXIS = 1   # y: the dam break is (nearly) uniform along y, so y-slabs stay balanced through the run
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # nominal non-tensor fp32 of a B200: 74.5
FLOP_PER_PARTICLE_PASS = 17.9e3    # ~767 candidates x 9 + ~137 neighbours x 80
print_json = lambda d: print(json.dumps(d), flush=True)


def workload_particles(world):
    return C3_PARTICLES if world <= 1 else int(C4_PARTICLES * world / 8)


def build_case(n_target, float_type="float32"):
    from sphexample_b200 import cases
    dp = cases.dp_for_count_3d(int(n_target))
    return cases.case_dam_break_3d(dp, float_type), dp


def params_of(case):
    from sphexample_b200 import make_params
    return make_params(case.meta, case.consts, case.kernel, case.viscosity, case.diffusion)


def host_threads():
    """host cores this process may use — NOT OMP_NUM_THREADS, which torchrun sets to 1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(n_particles, gpus, prep_time):
    """identical for both arms (the driver compares the two lines' config)"""
    name = "C3" if gpus <= 1 else ("C4" if gpus == 8 else f"C4 particle density on {gpus} of its 8 slabs")
    return {"workload": f"{name}: 3D dam-break lattice (example/Dambreak3d.jl constants, h=sqrt(3)dp, Wendland C2, artificial "
                        f"viscosity + linear density diffusion), state {prep_time:g} s after release "
                        f"({'developed flow' if prep_time > 0 else 'at rest'})",
            "particles": int(n_particles), "particles_per_gpu": int(round(n_particles / max(1, gpus))),
            "precision": "fp32 particle state",
            "parallelism": "single GPU" if gpus <= 1 else f"y-slab decomposition over {gpus} GPUs, NCCL halo exchange",
            "l2_policy": "L2 flushed between timed steps (256 MiB memset outside the per-step CUDA-event pairs); "
                         "value = particles*K / sum of per-step device times",
            "rebuild": "one forced neighbour rebuild at the start of the timed region + displacement-triggered ones "
                       "(reference cadence: forced rebuild per output interval of ~200 steps)"}


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


# ------------------------------------------------------------------------------------------------
# the workload state: the dam break `prep_time` seconds after release
# ------------------------------------------------------------------------------------------------
def developed_state(case, prep_time, device=0):
    """Run the case itself (untimed; part of building the synthetic input) on one GPU and return the
    particle table as a SimParticles in float64.  Needs the library; returns None without a GPU."""
    import torch
    if prep_time <= 0 or not torch.cuda.is_available():
        return None
    from sphexample_b200.preprocess import make_particles
    from sphexample_b200.simulation import Simulation
    sim = Simulation(params_of(case), device=device)
    sim.upload(case.particles)
    sim.SimulationLoop(prep_time)
    st = sim.download(order="id", fields=("Position", "Velocity", "Density", "Type", "ID", "GroupMarker"))
    sim.close()
    return make_particles(st["Position"].astype(np.float64), st["Density"].astype(np.float64), st["Type"], st["GroupMarker"],
                          st["ID"], velocity=st["Velocity"].astype(np.float64), dtype=np.float64, sort_by_id=False)


def cpu_rate(case, particles, threads, steps, warmup, budget_s):
    """the oracle (C++ restatement of the reference's multi-threaded algorithm, fp64) on the host cores.
    Times `steps` steps unless they would exceed budget_s, then as many as fit (>= 2).  -> (Mpu/s, steps timed, seconds)"""
    from oracle import oracle as orc
    o = orc.Oracle(params_of(case), particles, nthreads=threads)
    t0 = time.perf_counter()
    o.step(1, True)                        # incl. the first rebuild
    t_first = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        if time.perf_counter() - t0 > 0.25 * budget_s:
            break
        o.step(1, False)
    k = int(max(2, min(steps, (budget_s - (time.perf_counter() - t0)) / max(t_first, 1e-3))))
    t1 = time.perf_counter()
    o.step(k, False)
    secs = time.perf_counter() - t1
    n = len(particles)
    o.close()
    return n * k / secs / 1e6, k, secs


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (the oracle port: Julia is not in this image) on
    ALL host threads, from the same developed-flow state as the GPU arm.  N = 1: the full workload;
    larger N: a bounded sample (a smaller lattice of the same case at the same physical time)."""
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = host_threads()
    budget_s = float(os.environ.get("SPHB200_REF_BUDGET_S", "100"))
    from sphexample_b200 import cases
    n_target = int(args.particles) if args.particles else workload_particles(args.gpus)
    n_full = cases.dam_break_3d_count(cases.dp_for_count_3d(n_target))     # the lattice's actual particle count (as the GPU arm reports)
    n_case = min(n_target, int(os.environ.get("SPHB200_REF_MAX_PARTICLES", "2100000")))
    case, dp = build_case(n_case, "float64")
    state, how = None, "at rest (no GPU to develop the flow)"
    try:
        case32, _ = build_case(n_case, "float32")
        state = developed_state(case32, args.prep_time, device=int(os.environ.get("LOCAL_RANK", "0")))
        if state is not None:
            how = f"{args.prep_time:g} s after release (input state generated with libsphb200, untimed)"
    except Exception as ex:   # the reference arm must not depend on the product
        how = f"at rest (developing the flow failed: {ex})"
    particles = state if state is not None else case.particles
    value, k, secs = cpu_rate(case, particles, threads, args.steps, args.warmup, budget_s)
    n = len(particles)
    sample = (f"{k} of the {args.steps} steps, {n} particles (dp={dp:.6f})"
              + ("" if n_case == n_target else f" = a smaller lattice of the same case standing in for the {n_full}-particle workload")
              + f", state {how}, fp64, {secs:.1f} s on {threads} host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * n_full / (value * 1e6), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n_full, args.gpus, args.prep_time),
            "arm": "C++ restatement of the reference's multi-threaded CPU algorithm (oracle/sph_oracle.cpp; Julia unavailable), "
                   "all host threads; ms_per_step is the full workload at the measured rate",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print_json(line)


def cpu_baseline_sample(case64, state, prep_time, steps=20):
    """the oracle port on all host threads: the SAME workload and state, a bounded number of steps (~15 s)"""
    from oracle import oracle as orc
    orc.build()
    threads = host_threads()
    budget_s = float(os.environ.get("SPHB200_CPU_BUDGET_S", "15"))
    particles = state if state is not None else case64.particles
    value, k, secs = cpu_rate(case64, particles, threads, steps, 1, budget_s)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{k} steps (after a warm-up step incl. rebuild) of the same {len(particles)}-particle workload, state "
                      f"{prep_time:g} s after release, fp64, {secs:.1f} s on {threads} host threads"}


# ------------------------------------------------------------------------------------------------
# slab-vs-single-GPU self check (N >= 2)
# ------------------------------------------------------------------------------------------------
def slab_selfcheck(torch, rank, world, local_rank, per_rank=250_000, steps=60):
    """The slab decomposition against ONE GPU on the same case: a dam break of `per_rank` particles
    per rank with a smooth velocity field (so that cell rebuilds, list builds and migration happen
    within `steps` steps); rho, v, x compared by particle ID.  -> dict for the JSON line."""
    from sphexample_b200 import slab
    from sphexample_b200.simulation import Simulation
    dist = torch.distributed
    case, dp = build_case(per_rank * world, "float32")
    P = case.particles
    fl = (P.Type == 1)
    x = P.Position.astype(np.float64)
    P.Velocity[:, 0] = (1.5 * np.sin(3.0 * x[:, 2] + 1.0) * fl).astype(np.float32)
    P.Velocity[:, 1] = (1.0 * np.sin(9.0 * x[:, 0] + 4.0 * x[:, 1]) * fl).astype(np.float32)
    P.Velocity[:, 2] = (-1.5 * np.cos(2.0 * x[:, 0]) * fl).astype(np.float32)
    p = params_of(case)
    sim = Simulation(p, device=local_rank)
    dec = slab.SlabDecomposition(sim, P, p.H_inv, rank, world, axis=SLAB_AXIS)
    dec.setup()
    rep = sim.step(steps, reset_delta_x=True)
    migrated = sim.stat("migrated")
    got = dec.gather(order="id", fields=("Position", "Velocity", "Density", "ID"))
    sim.close()
    mig = torch.tensor([float(migrated)], device="cuda", dtype=torch.float64)
    dist.all_reduce(mig)
    out = None
    if rank == 0:
        one = Simulation(p, device=local_rank)
        one.upload(P)
        rep1 = one.step(steps, reset_delta_x=True)
        ref = one.download(order="id", fields=("Position", "Velocity", "Density", "ID"))
        one.close()
        assert np.array_equal(ref["ID"], got["ID"])

        def rel(a, b):
            a, b = a.astype(np.float64), b.astype(np.float64)
            return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        out = {"particles": int(len(P)), "particles_per_rank": int(per_rank), "steps": int(steps),
               "cell_rebuilds": int(rep["n_rebuilds"]), "cell_rebuilds_single_gpu": int(rep1["n_rebuilds"]),
               "migrated_particles": int(mig.item()),
               "max_rel_err_vs_single_gpu": {"Density": rel(got["Density"], ref["Density"]), "Velocity": rel(got["Velocity"], ref["Velocity"]),
                                             "Position": rel(got["Position"], ref["Position"])},
               "total_time_equal": bool(rep["total_time"] == rep1["total_time"]),
               "note": "fp32; the slab run sums the same pairs in another order (y-major cell key), so fp32 rounding differs"}
    dist.barrier()
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    from sphexample_b200.simulation import Simulation
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libsphb200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = torch.distributed if world > 1 else None
    n_total = int(args.particles) if args.particles else workload_particles(world)
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
        # slab edges: minimise the largest slab load; a wall particle counts 0.8 of a fluid particle — measured
        # on C4 (profiles/r3d_bench_8gpu_c4.json, per_rank): the two end ranks, which own the side walls, spend
        # 0.42 us per owned particle and pass against 0.45 for the interior ranks
        dec = slab.SlabDecomposition(sim, parts, p.H_inv, rank, world, axis=SLAB_AXIS,
                                     boundary_weight=float(os.environ.get("SPHB200_BOUNDARY_WEIGHT", "0.8")))
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
    st0 = sim.download(fields=("Position", "Velocity", "Density", "Type", "ID", "GroupMarker"))
    vmax_local = float(np.sqrt((st0["Velocity"].astype(np.float64) ** 2).sum(1)).max())
    host = {k: torch.from_numpy(np.ascontiguousarray(st0[k])).pin_memory().numpy() for k in ("Position", "Velocity", "Density")}
    # (every column of the host table is pinned: a pageable column would go through the driver's staging buffer)
    types = torch.from_numpy(np.ascontiguousarray(st0["Type"], np.uint8)).pin_memory().numpy()
    ids = torch.from_numpy(np.ascontiguousarray(st0["ID"], np.int64)).pin_memory().numpy()
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

    prep["vmax"] = max_over_ranks([vmax_local])[0]
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
    step_ms = [float(a.elapsed_time(b)) for a, b in ev]
    ms = float(sum(step_ms))
    srt = sorted(step_ms)
    # where the K per-step times lie (this rank): a cell rebuild + full list build shows as a few slow steps
    step_spread = {"min": round(srt[0], 4), "median": round(srt[len(srt) // 2], 4), "max": round(srt[-1], 4),
                   "slowest": [[int(i), round(step_ms[i], 4)] for i in sorted(range(len(step_ms)), key=lambda i: -step_ms[i])[:3]]}
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

    # ---- per-stage device times (CUDA events on the library's stream, extra steps, L2 flushed) ----
    # every stage: mean over `reps` sampled steps (so the rare cell rebuild and its full list build appear at
    # their amortised cost per step); the two pass stages used for the roofline: median of the samples
    reps = 48
    samples = []
    for _ in range(reps):
        flush.zero_()
        if world > 1:
            dist.barrier()
        samples.append(sim.stage_times())
    names = list(samples[0].keys())
    mat = np.array([[s[k] for k in names] for s in samples])
    stage_mean = np.array(max_over_ranks(mat.mean(0)))
    stage_ms = {k: float(v) for k, v in zip(names, stage_mean)}
    k1 = names.index("05-07 First NeighborLoop + half step (fused)")
    k2 = names.index("08-11 Second NeighborLoop + full step (fused)")
    pass1_ms, pass2_ms = float(np.median(mat[:, k1])), float(np.median(mat[:, k2]))
    per_rank = None
    if world > 1:   # what every rank spends in its passes and owns: the load balance of the slab decomposition
        t = torch.tensor([pass1_ms, pass2_ms, float(sim.report()["n_particles"]), float(sim.stat("list_entries"))],
                         device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allr, t)
        per_rank = {"pass1_ms": [round(float(a[0]), 4) for a in allr], "pass2_ms": [round(float(a[1]), 4) for a in allr],
                    "owned": [int(a[2]) for a in allr], "list_entries_per_particle": [round(float(a[3]), 1) for a in allr]}
    pass1_ms, pass2_ms = max_over_ranks([pass1_ms, pass2_ms])
    n_local_max = int(max_over_ranks([sim.report()["n_particles"]])[0])
    list_entries = max_over_ranks([sim.stat("list_entries")])[0]
    list_wavefronts = max_over_ranks([sim.stat("list_wavefronts")])[0]
    halo_bytes = sum_over_ranks(int(sim.stat("halo_bytes_per_step"))) if world > 1 else 0
    D, sz = 3, 4
    bytes_pass0 = n_local_max * (4 * D + 4) * sz          # read x,v,rho,P ; write x_h,v_h,rho_h,P_h
    bytes_pass1 = n_local_max * (7 * D + 5) * sz          # read half state + own state n + rho_n ; write x,v,rho,P,a
    peak, how = measured_peaks()
    t_avg = 0.5 * (pass1_ms + pass2_ms) * 1e-3
    achieved = 0.5 * (bytes_pass0 + bytes_pass1) / t_avg / 1e9
    prof = os.path.join(ROOT, "profiles", "interact_traffic.json")
    traffic, traffic_src = None, None
    if os.path.exists(prof):
        tj = json.load(open(prof))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    roof = {"bound": "hbm", "kernel": "k_interact_ring<float,3,PASS> (list kernel: TMA producer warp + 8 consumer warps per SM); median of "
                              "each pass over the sampled steps, mean of the two passes, slowest rank",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "peak_source": how,
            "algorithmic_bytes_per_launch": 0.5 * (bytes_pass0 + bytes_pass1), "avg_launch_ms": 0.5 * (pass1_ms + pass2_ms),
            "pass1_ms": pass1_ms, "pass2_ms": pass2_ms,
            "note": "the interaction kernel is NOT HBM-bound (~18 kflop per particle per pass, SURVEY 8d): it is bound by fp32 issue "
                    "slots; see roofline_fp32 for the binding ceiling.  The HBM fraction is reported because the metric asks for it."}
    flops = FLOP_PER_PARTICLE_PASS * n_local_max
    roof32 = {"bound": "fp32 (non-tensor FMA pipe; no dense contraction on this path)", "achieved": flops / t_avg / 1e12,
              "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": flops / t_avg / 1e12 / FP32_PEAK_TFLOPS,
              "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (SURVEY 8d); no measured non-tensor fp32 peak in MEASURED_PEAKS.json",
              "algorithmic_flop_per_particle_pass": FLOP_PER_PARTICLE_PASS,
              "list_entries_per_particle": list_entries, "list_smem_wavefronts_per_quarter_warp_gather": list_wavefronts}

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
    k2r = 5
    barrier()
    e0.record(stream)
    for _ in range(k2r):
        upload()
        sim.step(1, reset_delta_x=True)
        sim.download_into(out["Position"][:sim.num_particles], out["Velocity"][:sim.num_particles],
                          out["Density"][:sim.num_particles], out["Pressure"][:sim.num_particles])
    e1.record(stream)
    barrier()
    e2e["roundtrip_every_step_value"] = n * k2r / (max_over_ranks([e0.elapsed_time(e1)])[0] * 1e-3) / 1e6
    sim.close()

    selfcheck = None
    if world > 1 and not args.no_selfcheck:
        try:
            selfcheck = slab_selfcheck(torch, rank, world, local_rank)
        except Exception as ex:   # a failing check is reported, it must not void the timing
            selfcheck = {"failed": repr(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(n, world, args.prep_time), "clocks": clocks,
                "gpu_launches": int(launches), "roofline": roof, "roofline_fp32": roof32, "e2e": e2e,
                "stage_ms": stage_ms,
                "rebuilds_in_timed_region": rebuilds_timed, "list_builds_in_timed_region": list_builds_timed, "dp": dp,
                "step_ms_spread_rank0": step_spread,
                "stage_ms_note": "stage_ms samples the FULL step sequence (sphb200_stage_times: UpdateNeighbors! chain and stand-by "
                                 "kernels in place, running empty when not needed); on one GPU the timed steps replay the lean "
                                 "sequence, which has neither, so the stages add up to slightly more than ms_per_step",
                "state": prep,
                "back_to_back": {"value": n * args.steps / (ms_b2b * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_b2b / args.steps,
                                 "note": "same K steps enqueued back to back, warm L2"}}
        if world > 1:
            line["scaling_note"] = (f"weak scaling at {C4_PARTICLES // 8} particles per GPU for N >= 2 (N = 8 is C4); the N = 1 line is C3 "
                                    f"({C3_PARTICLES} particles), the configuration the metric is quoted on")
            line["slab"] = {"axis": "xyz"[SLAB_AXIS], "edges": [int(e) for e in dec.edges], "boundary_weight": dec.boundary_weight,
                            "owned_max": n_local_max, "owned_mean": n / world, "halo_bytes_per_step": halo_bytes, "per_rank": per_rank}
            line["selfcheck"] = selfcheck
        # ---- CPU baseline (oracle port) on a bounded sample of the same workload, N = 1 only ----
        if world == 1:
            try:
                from sphexample_b200.preprocess import make_particles
                case64, _ = build_case(n_total, "float64")
                state = make_particles(st0["Position"].astype(np.float64), st0["Density"].astype(np.float64), st0["Type"],
                                       st0["GroupMarker"], st0["ID"], velocity=st0["Velocity"].astype(np.float64),
                                       dtype=np.float64, sort_by_id=True)
                line["cpu_baseline"] = cpu_baseline_sample(case64, state, args.prep_time)
            except Exception as ex:   # the checker failing must not void the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
        print_json(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=float, default=0, help="override the total particle count")
    ap.add_argument("--prep-time", type=float, default=0.15,
                    help="simulated seconds the dam break runs (untimed) before the benchmark state is taken; 0 = from rest")
    ap.add_argument("--no-selfcheck", action="store_true", help="N >= 2: skip the slab-vs-single-GPU comparison")
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
    # stdout carries exactly ONE line, the JSON: anything a library prints there (NCCL's version banner,
    # for one) is sent to stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global print_json
    print_json = lambda d: (json_out.write(json.dumps(d) + "\n"), json_out.flush())
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
