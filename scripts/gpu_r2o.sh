#!/bin/bash
# round 2, call o: where does the per-brick list maintenance spend its time (launch list, developed flow)
O=gpurun_out/r2o; mkdir -p $O
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 12 > $O/launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, io, collections
rows = list(csv.DictReader(io.StringIO("".join(l for l in open("gpurun_out/r2o/launches.csv") if l.startswith('"')))))
agg = collections.defaultdict(list)
for r in rows:
    agg[r["Kernel Name"].split("(")[0][:44]].append(float(r["Metric Value"].replace(",", "")) / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:9]:
    print(f"{k:46s} n={len(v):3d} sum={sum(v):9.1f}us max={max(v):8.1f}us min={min(v):8.1f}us")
PY
