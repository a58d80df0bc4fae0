#!/bin/bash
# lean step sequence + warp-cooperative mDBC gather: tests, then the small-config anatomy again
O=gpurun_out/r4c; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_lean.py tests/test_gpu_slab.py tests/test_gpu_parity.py -q -m gpu -x -k "lean or mdbc or simulation_loop or world_of_one" > $O/pytest_lean.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_lean.log
timeout 200 python scripts/small_profile.py > $O/small_profile.jsonl 2> $O/small_profile.err; echo "profile rc=$?"; cut -c1-330 $O/small_profile.jsonl; tail -3 $O/small_profile.err
OPT_LIST_LOCAL=0 timeout 200 python scripts/small_profile.py > $O/small_profile_global_lists.jsonl 2>> $O/small_profile.err; echo "profile2 rc=$?"; cut -c1-200 $O/small_profile_global_lists.jsonl
OPT_LEAN=0 timeout 200 python scripts/small_profile.py c1 c5 > $O/small_profile_nolean.jsonl 2>> $O/small_profile.err; echo "profile3 rc=$?"; cut -c1-200 $O/small_profile_nolean.jsonl
