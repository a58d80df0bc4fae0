#!/bin/bash
# round 2, final single-GPU evidence of the last build: smoke, full GPU suite with margins, bench at the driver's K and
# at 200, reference arm, config table, launch list + ncu --set full of the final build, sanitizer
O=gpurun_out/r4z; mkdir -p $O
(nvidia-smi; nproc; lscpu | head -20) > $O/host.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
rm -f $O/parity.jsonl
SPH_PARITY_LOG=$PWD/$O/parity.jsonl timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_k20.json 2> $O/bench_k20.err; echo "bench k20 rc=$?"; tail -2 $O/bench_k20.err | cut -c1-300
timeout 600 python bench.py --steps 200 --warmup 20 > $O/bench.json 2> $O/bench.err; echo "bench k200 rc=$?"; tail -2 $O/bench.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for k in ("bench_k20", "bench"):
    d = json.loads([l for l in open(f"gpurun_out/r4z/{k}.json") if l.startswith("{")][0])
    print(k, {x: d[x] for x in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "b2b", d["back_to_back"]["value"],
          "frac", d["roofline"]["frac"], d["roofline_fp32"]["frac"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for n, v in d["stage_ms"].items():
    if v: print(f"  {v:8.4f}  {n}")
r = json.loads([l for l in open("gpurun_out/r4z/bench_ref.json") if l.startswith("{")][0])
print("ref", r["value"], r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"])
PY
timeout 600 python scripts/config_table.py > $O/configs.jsonl 2> $O/configs.err; echo "configs rc=$?"; cut -c1-260 $O/configs.jsonl
SPH_PREP=0.15 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 24 > $O/launches.log 2>&1; echo "launch list rc=$?"
SPH_PREP=0.15 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_interact_ring" -c 2 -f -o $O/prof_interact python scripts/profile_step.py 1e6 4 > $O/prof.log 2>&1; echo "ncu rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_case.py 6 > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 $O/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_case.py 3 > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -1 $O/sanitizer_racecheck.log
