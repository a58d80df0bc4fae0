#!/bin/bash
# round 2, call s: upload/download without host passes; GPU suite; bench N=1 at the driver's K (20) and at 200; ncu of the build
O=gpurun_out/r2s; mkdir -p $O
(nvidia-smi; nproc; lscpu | head -20) > $O/host.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench k20 rc=$?"; tail -2 $O/bench_k20.err
timeout 900 python bench.py --steps 200 --warmup 20 > $O/bench_k200.json 2> $O/bench_k200.err; echo "bench k200 rc=$?"; tail -2 $O/bench_k200.err
python - <<'PY'
import json
for k in ("k20", "k200"):
    d = json.loads([l for l in open(f"gpurun_out/r2s/bench_{k}.json") if l.startswith("{")][0])
    print(k, {x: d[x] for x in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "b2b", d["back_to_back"]["value"],
          "frac", d["roofline"]["frac"], d["roofline_fp32"]["frac"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    if k == "k200":
        for n, v in d["stage_ms"].items(): print(f"  {v:8.4f}  {n}")
PY
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 24 > $O/launches.log 2>&1; echo "launch list rc=$?"
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_interact_ring" -c 2 -f -o $O/prof_interact python scripts/profile_step.py 1e6 4 > $O/prof.log 2>&1; echo "ncu rc=$?"; tail -1 $O/prof.log | cut -c1-100
