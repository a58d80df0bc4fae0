#!/bin/bash
# round 2, call b: first run of the ring kernel — parity (lists / parity suites), sweep, launch list, ncu of the ring kernel
O=gpurun_out/r2b; mkdir -p $O
timeout 60 scripts/ubench/ubench.bin > $O/ubench.jsonl 2>&1; echo "ubench rc=$?"; cat $O/ubench.jsonl
timeout 900 python -m pytest tests/test_gpu_lists.py tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_lists_parity.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_lists_parity.log
SPH_SWEEP="lists=1;lists=1,list_reorder=0;lists=1,skin=0.07;lists=1,skin=0.15;lists=0" SPH_STEPS=120 timeout 600 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cat $O/tune.jsonl; tail -3 $O/tune.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 4 > $O/launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_interact_ring|k_list_reorder|k_list_build" -s 6 -c 4 -f -o $O/prof_interact python scripts/profile_step.py 1e6 2 > $O/prof.log 2>&1; echo "ncu rc=$?"; tail -2 $O/prof.log
