#!/bin/bash
# bricks sized by particle count (32-target bricks for the small cases): full GPU suite, small configs
O=gpurun_out/r4i; mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 200 python scripts/small_profile.py > $O/small_profile.jsonl 2> $O/small_profile.err; echo "profile rc=$?"; cut -c1-200 $O/small_profile.jsonl; tail -3 $O/small_profile.err
OPT_BRICK_TARGETS=128 timeout 200 python scripts/small_profile.py c1 c2 > $O/small_profile_bt128.jsonl 2>> $O/small_profile.err; cut -c1-200 $O/small_profile_bt128.jsonl
OPT_BRICK_TARGETS=64 timeout 200 python scripts/small_profile.py c1 c2 > $O/small_profile_bt64.jsonl 2>> $O/small_profile.err; cut -c1-200 $O/small_profile_bt64.jsonl
