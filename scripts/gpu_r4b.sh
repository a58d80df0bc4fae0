#!/bin/bash
# small-config step anatomy + the RunSimulation/VTKHDF GPU test
O=gpurun_out/r4b; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_runsimulation.py -q -m gpu -x > $O/pytest_runsim.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_runsim.log
timeout 200 python scripts/small_profile.py > $O/small_profile.jsonl 2> $O/small_profile.err; echo "profile rc=$?"; cat $O/small_profile.jsonl; tail -3 $O/small_profile.err
SMALL_PROFILE_NCU=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/small_launches.csv python scripts/small_profile.py c1 c5 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r4b/small_launches.csv") if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki].split("(")[0][:60], float(r[vi].replace(",", ""))) for r in rows[1:]]
print(len(seq), "launches")
agg = collections.OrderedDict()
for k, v in seq[-600:]:
    agg.setdefault(k, []).append(v)
for k, v in agg.items():
    v.sort(); print(f"{k:62s} n={len(v):4d} median={v[len(v)//2]/1e3:8.2f} us max={v[-1]/1e3:8.2f}")
PY
