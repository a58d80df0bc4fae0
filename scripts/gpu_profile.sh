#!/bin/bash
# ncu launch list + full capture of the interaction kernel + option sweep
mkdir -p gpurun_out
TAG=${1:-r1a}
timeout 600 python scripts/sweep.py > gpurun_out/sweep_$TAG.jsonl 2> gpurun_out/sweep_$TAG.err; echo "sweep rc=$?"; cat gpurun_out/sweep_$TAG.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_step.py 1e6 2 > gpurun_out/launches_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_interact -s 6 -c 2 -f -o gpurun_out/prof_interact_$TAG python scripts/profile_step.py 1e6 1 > gpurun_out/prof_$TAG.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/prof_$TAG.log
ls -la gpurun_out/
