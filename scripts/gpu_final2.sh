#!/bin/bash
# final checks of the last commit: build(), smoke(), full GPU suite, bench N=1 (K=20 and default)
O=gpurun_out/r3c; mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 2400 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench k20 rc=$?"; tail -2 $O/bench_k20.err
/usr/bin/time -v timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench default rc=$?"; grep -E "Elapsed" $O/bench.err
python - <<'PY'
import json
for k in ("bench_k20", "bench"):
    d = json.loads([l for l in open(f"gpurun_out/r3c/{k}.json") if l.startswith("{")][0])
    print(k, {x: d[x] for x in ("value", "ms_per_step", "gpu_launches", "steps")}, "e2e", d["e2e"]["value"], "b2b", d["back_to_back"]["value"],
          "frac", round(d["roofline"]["frac"], 4), round(d["roofline_fp32"]["frac"], 3), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
PY
