#!/bin/bash
# One GPU visit: smoke, parity suite, bench (ours + reference arm), ncu launch list and a full
# capture of the interaction kernel.  Everything is logged under gpurun_out/<TAG>/.
TAG=${1:-r1}
O=gpurun_out/$TAG; mkdir -p $O
nproc > $O/host.txt; lscpu | head -25 >> $O/host.txt; nvidia-smi >> $O/host.txt 2>&1
export SPH_PARITY_LOG=$PWD/$O/parity.jsonl; rm -f $SPH_PARITY_LOG
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 10 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cat $O/bench.json; tail -5 $O/bench.err
echo "== bench reference arm"; SPHB200_REF_BUDGET_S=40 timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err; cat $O/bench_ref.json
echo "== configs"; timeout 1200 python scripts/config_table.py > $O/configs.jsonl 2> $O/configs.err; echo "configs rc=$?"; cat $O/configs.jsonl; tail -3 $O/configs.err
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 4 > $O/launches.log 2>&1; echo "launch list rc=$?"
echo "== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_interact|k_list_build" -s 8 -c 4 -f -o $O/prof_interact python scripts/profile_step.py 1e6 2 > $O/prof.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/prof.log
ls -la $O
