#!/bin/bash
# round 2, call g: packed (FFMA2) pair body in the ring kernel
O=gpurun_out/r2g; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_lists.py tests/test_gpu_parity.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
SPH_SWEEP="lists=1" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 12 > $O/launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, io, collections
rows = list(csv.DictReader(io.StringIO("".join(l for l in open("gpurun_out/r2g/launches.csv") if l.startswith('"')))))
agg = collections.defaultdict(list)
for r in rows:
    agg[r["Kernel Name"].split("(")[0][:44]].append(float(r["Metric Value"].replace(",", "")) / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:6]:
    print(f"{k:46s} n={len(v):3d} sum={sum(v):9.1f}us max={max(v):8.1f}us")
PY
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_interact_ring" -c 2 -f -o $O/prof_interact python scripts/profile_step.py 1e6 4 > $O/prof.log 2>&1; echo "ncu rc=$?"; tail -1 $O/prof.log | cut -c1-100
