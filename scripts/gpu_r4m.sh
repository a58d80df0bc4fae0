#!/bin/bash
# last build (lean estimate moved into sph_control.h): smoke + full GPU suite + a short C3 bench
O=gpurun_out/r4m; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 700 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r4m/bench_k20.json") if l.startswith("{")][0])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "b2b", d["back_to_back"]["value"], d["step_ms_spread_rank0"], "cpu", d["cpu_baseline"]["value"])
PY
