#!/bin/bash
# round 2, call f: build walk split by row role; NCW variants; skin sweep
O=gpurun_out/r2f; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_lists.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
SPH_SWEEP="lists=1;lists=1,skin=0.07;lists=1,skin=0.13" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
for v in ncw12 ncw10; do
  SPHB200_LIB=sphexample_b200/lib/libsphb200_$v.so SPH_SWEEP="lists=1" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune_$v.jsonl 2> $O/tune_$v.err; echo "tune $v rc=$?"; cut -c1-330 $O/tune_$v.jsonl; tail -2 $O/tune_$v.err
done
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 12 > $O/launches.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv, io, collections
rows = list(csv.DictReader(io.StringIO("".join(l for l in open("gpurun_out/r2f/launches.csv") if l.startswith('"')))))
agg = collections.defaultdict(list)
for r in rows:
    agg[r["Kernel Name"].split("(")[0][:44]].append(float(r["Metric Value"].replace(",", "")) / 1e3)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:6]:
    print(f"{k:46s} n={len(v):3d} sum={sum(v):9.1f}us max={max(v):8.1f}us")
PY
