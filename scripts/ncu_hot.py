"""Hottest instructions (stall samples) of one kernel launch in an .ncu-rep, with the SASS around them.
  python scripts/ncu_hot.py report.ncu-rep <kernel regex> [launch index among matches] [min pct] [context]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
ctx = int(sys.argv[5]) if len(sys.argv) > 5 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
k, hdr, out, name = -1, None, [], ""
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Kernel Name":
        k += 1
        if k == which: name = r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if k != which or hdr is None or len(r) < len(hdr) - 5: continue
    out.append(dict(zip(hdr, r)))
tot = sum(int(d["# Samples"]) for d in out) or 1
print(name[:110]); print("samples", tot, "instructions", len(out))
keys = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math", "stall_mio", "stall_not_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst"]
agg = {k2: sum(int(d.get(k2, "0") or 0) for d in out) for k2 in keys}
print("stall mix:", {k2[6:]: round(v / tot * 100, 1) for k2, v in agg.items()})
hot = [n for n, d in enumerate(out) if int(d["# Samples"]) / tot * 100 >= minpct]
shown = set()
for n in hot:
    for m in range(max(0, n - ctx), min(len(out), n + ctx + 1)):
        if m in shown: continue
        shown.add(m)
        d = out[m]
        top = max(keys, key=lambda k2: int(d.get(k2, "0") or 0))
        print(f"{m:5d} {int(d['# Samples']) / tot * 100:5.1f}% {top[6:]:>12s} exec {d['Instructions Executed']:>9s}  {d['Source'].strip()[:90]}")
    if ctx: print("  ...")
