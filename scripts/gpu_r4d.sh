#!/bin/bash
# split lists (fp64 2D, small N) + list_local by size + graphs kept across same-size uploads: full GPU suite, small configs, C3 bench
O=gpurun_out/r4d; mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 200 python scripts/small_profile.py > $O/small_profile.jsonl 2> $O/small_profile.err; echo "profile rc=$?"; cut -c1-200 $O/small_profile.jsonl; tail -3 $O/small_profile.err
OPT_SPLIT=0 timeout 200 python scripts/small_profile.py c1 c5 > $O/small_profile_nosplit.jsonl 2>> $O/small_profile.err; echo "profile2 rc=$?"; cut -c1-200 $O/small_profile_nosplit.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r4d/bench.json") if l.startswith("{")][0])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "b2b", d["back_to_back"]["value"], "cpu", d["cpu_baseline"]["value"])
print({k: round(v, 4) for k, v in d["stage_ms"].items() if v})
PY
