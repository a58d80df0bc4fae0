#!/bin/bash
# round 2, call d: ring kernel after the write-after-read fix of the list prefetch; ncu of real build / reorder launches
O=gpurun_out/r2d; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_lists.py -x -q -m gpu > $O/pytest_lists.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_lists.log
SPH_SWEEP="lists=1;lists=1,list_reorder=0" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cat $O/tune.jsonl; tail -3 $O/tune.err
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 12 > $O/launches.log 2>&1; echo "launch list rc=$?"; tail -1 $O/launches.log
python - <<'PY'
import csv, io
rows = list(csv.DictReader(io.StringIO("".join(l for l in open("gpurun_out/r2d/launches.csv") if l.startswith('"')))))
for r in rows:
    n = r["Kernel Name"]
    if "list_build" in n or "list_reorder" in n or "ring" in n:
        print(n[:40], r["Metric Value"])
PY
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_interact_ring" -c 2 -f -o $O/prof_interact python scripts/profile_step.py 1e6 4 > $O/prof.log 2>&1; echo "ncu rc=$?"; tail -2 $O/prof.log
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_list_reorder|k_list_build" -c 16 -f -o $O/prof_build python scripts/profile_step.py 1e6 8 > $O/prof2.log 2>&1; echo "ncu2 rc=$?"; tail -2 $O/prof2.log
