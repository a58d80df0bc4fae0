#!/bin/bash
# round 2, call r: conditional-node step graph — GPU suite, config table (C1/C2/C5), C3 timing
O=gpurun_out/r2r; mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu -x --durations=4 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -10 $O/pytest_gpu.log
timeout 900 python scripts/config_table.py > $O/configs.jsonl 2> $O/configs.err; echo "configs rc=$?"; python - <<'PY'
import json
for l in open("gpurun_out/r2r/configs.jsonl"):
    d = json.loads(l)
    print(f"{d['gpu_Mpu_s']:9.1f} Mpu/s  cpu1 {d['cpu_1thread_Mpu_s']:7.3f} cpuN {d['cpu_all_Mpu_s']:7.3f}  err v {d['err_vel']:.1e} rho {d['err_rho']:.1e}  rebuilds {d['rebuilds']} lists {d['list_builds']:.1f}  {d['config']}")
PY
tail -3 $O/configs.err
SPH_SWEEP="lists=1;lists=1,graph_cond=0" SPH_STEPS=200 timeout 400 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
