"""Per-source-line instruction shares of one kernel launch from an .ncu-rep (needs -lineinfo).
  python scripts/ncu_lines.py report.ncu-rep [launch_index] [min_pct]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
sections = []; cur = None
for r in rows:
    if r and r[0] == "File Path": cur = {"file": r[1], "rows": []}; sections.append(cur)
    elif r and r[0] == "Function Name": cur["func"] = r[1]
    elif r and r[0] == "Line No": cur["hdr"] = r
    elif cur is not None and r: cur["rows"].append(r)
funcs = []
for s in sections:
    if s["func"] not in funcs: funcs.append(s["func"])
fn = funcs[which]
agg = collections.Counter(); samp = collections.Counter(); thr = collections.Counter(); src = {}; tot = 0
for s in sections:
    if s["func"] != fn: continue
    h = s["hdr"]; iI = h.index("Instructions Executed"); iN = h.index("# Samples"); iT = h.index("Thread Instructions Executed")
    f = s["file"].split("/")[-1]
    for r in s["rows"]:
        if not r[0].strip(): continue
        try: n = int(r[iI])
        except ValueError: continue
        k = (f, int(r[0])); agg[k] += n; tot += n
        try: samp[k] += int(r[iN]); thr[k] += int(r[iT])
        except ValueError: pass
        src.setdefault(k, r[1].strip()[:95])
print(fn[:100]); print("total warp instructions", tot)
ts = sum(samp.values())
for k, v in sorted(agg.items()):
    if v / tot * 100 < minpct: continue
    print(f"{v/tot*100:5.1f}% inst {samp[k]/max(ts,1)*100:5.1f}% samp lanes {thr[k]/max(v,1):4.1f}  {k[0][:18]}:{k[1]:<4d} {src.get(k,'')}")
