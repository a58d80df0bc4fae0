#!/bin/bash
# round 2, call c: ring kernel with unconditional list prefetch; NCW / NSLOT variants; developed-flow launch list + ncu
O=gpurun_out/r2c; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_lists.py -x -q -m gpu > $O/pytest_lists.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_lists.log
SPH_SWEEP="lists=1;lists=1,list_reorder=0" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cat $O/tune.jsonl; tail -3 $O/tune.err
for v in ncw12 ncw6 ncw12s4; do
  SPHB200_LIB=sphexample_b200/lib/libsphb200_$v.so SPH_SWEEP="lists=1" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune_$v.jsonl 2> $O/tune_$v.err; echo "tune $v rc=$?"; cat $O/tune_$v.jsonl; tail -2 $O/tune_$v.err
done
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python scripts/profile_step.py 1e6 12 > $O/launches.log 2>&1; echo "launch list rc=$?"; tail -1 $O/launches.log
SPH_PREP=0.15 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_interact_ring|k_list_reorder|k_list_build" -c 6 -f -o $O/prof_interact python scripts/profile_step.py 1e6 8 > $O/prof.log 2>&1; echo "ncu rc=$?"; tail -2 $O/prof.log
