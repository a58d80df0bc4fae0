#!/bin/bash
# round 2, call v: rho_n snapshot folded into the pass-1 epilogue; warp-per-cell velocity boxes
O=gpurun_out/r2v; mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
SPH_SWEEP="lists=1" SPH_STEPS=200 timeout 400 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 24 > $O/launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, io, collections
rows = list(csv.DictReader(io.StringIO("".join(l for l in open("gpurun_out/r2v/launches.csv") if l.startswith('"')))))
agg = collections.defaultdict(list)
for r in rows:
    agg[r["Kernel Name"].split("(")[0][:44]].append(float(r["Metric Value"].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
print("total per step us", tot / 24)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:12]:
    print(f"{k:46s} n={len(v):3d} per-step={sum(v)/24:8.1f}us max={max(v):8.1f}us")
PY
