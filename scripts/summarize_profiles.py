"""Turn one gpurun_out/<tag>/ directory into the committed summaries under profiles/<tag>_*.
  python scripts/summarize_profiles.py gpurun_out/r1b r1b
Reads: launches.csv (ncu launch list), prof_interact.ncu-rep (ncu --set full), bench.json,
parity.jsonl, sweep.jsonl, host.txt — whichever exist."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

src, tag = sys.argv[1], sys.argv[2]
os.makedirs("profiles", exist_ok=True)
P = lambda name: os.path.join("profiles", f"{tag}_{name}")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def launch_list(path):
    text = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(text))))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("sph::", "")
        if "k_interact" in r["Kernel Name"]:
            k = r["Kernel Name"].split("(sph::")[0].replace("void ", "").replace("sph::", "").replace("(int)", "").replace("(bool)", "")
        v = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list ({len(rows)} launches, "
           f"total {tot / 1e6:.3f} ms; per-launch times are cold-cache and serialised: compare shares)",
           f"{'share':>7s} {'count':>6s} {'avg_us':>10s}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1] / tot * 100:6.2f}% {v[0]:6d} {v[1] / v[0] / 1e3:10.2f}  {k}")
    return "\n".join(out) + "\n"


def full_capture(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out, js = [], []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        out.append(f"## {name}")
        d = {"kernel": name}
        for m in METRICS:
            if m in idx:
                out.append(f"{m:70s} {r[idx[m]]:>16s} {units[idx[m]]}")
                try:
                    d[m] = float(r[idx[m]].replace(",", ""))
                except ValueError:
                    d[m] = r[idx[m]]
                d[m + "__unit"] = units[idx[m]]
        js.append(d)
        out.append("")
    return "\n".join(out) + "\n", js


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


if os.path.exists(f"{src}/launches.csv"):
    open(P("launches.txt"), "w").write(launch_list(f"{src}/launches.csv"))
for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
    base = rep[:-8]
    txt, js = full_capture(f"{src}/{rep}")
    open(P(f"{base}_ncu_full.txt"), "w").write(f"# ncu --set full --clock-control none --import-source on, {rep}\n" + txt)
    if "interact" in base and js:
        # launches that did the work (the predicated twins of a pass return in microseconds)
        def dur_us(d):
            return d["gpu__time_duration.sum"] * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["gpu__time_duration.sum__unit"], 1.0)
        busy = [d for d in js if dur_us(d) > 50.0 and "k_list_build" not in d["kernel"]]
        tr = [to_bytes(d["dram__bytes_read.sum"], d["dram__bytes_read.sum__unit"]) +
              to_bytes(d["dram__bytes_write.sum"], d["dram__bytes_write.sum__unit"]) for d in busy]
        if tr:
            json.dump({"source": f"profiles/{tag}_{base}_ncu_full.txt", "kernels": [d["kernel"] for d in busy],
                       "duration_us": [dur_us(d) for d in busy],
                       "dram_bytes_per_launch_each": tr, "dram_bytes_per_launch": sum(tr) / len(tr)},
                      open("profiles/interact_traffic.json", "w"), indent=1)
    lines = subprocess.run([sys.executable, "scripts/ncu_lines.py", f"{src}/{rep}", "0", "0.8"], capture_output=True, text=True).stdout
    if lines.strip():
        open(P(f"{base}_source_lines.txt"), "w").write(lines)
for name in ("bench.json", "configs.jsonl", "sweep.jsonl", "host.txt", "smoke.log", "pytest_gpu.log", "bench_ref.json", "slab.jsonl"):
    if os.path.exists(f"{src}/{name}"):
        open(P(name), "w").write(open(f"{src}/{name}").read())
if os.path.exists(f"{src}/parity.jsonl"):
    rows = [json.loads(l) for l in open(f"{src}/parity.jsonl")]
    worst = {}
    for r in rows:
        t = r["test"].split("::")[1]
        w = worst.get(t)
        if w is None or r["err"] / r["tol"] > w["err"] / w["tol"]:
            worst[t] = r
    with open(P("parity.txt"), "w") as f:
        f.write(f"# GPU parity margins vs the CPU oracle ({len(rows)} checks, all ok = {all(r['ok'] for r in rows)}); worst check per test\n")
        for t, r in sorted(worst.items()):
            f.write(f"{r['err']:10.3e} (tol {r['tol']:.0e})  {t}\n")
print("wrote", sorted(f for f in os.listdir("profiles") if f.startswith(tag) or f == "interact_traffic.json"))
