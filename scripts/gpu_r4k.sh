#!/bin/bash
# 2D fp64 list kernel with 11 consumer warps: full GPU suite, small configs
O=gpurun_out/r4k; mkdir -p $O
timeout 700 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 200 python scripts/small_profile.py > $O/small_profile.jsonl 2> $O/small_profile.err; echo "profile rc=$?"; cut -c1-200 $O/small_profile.jsonl; tail -3 $O/small_profile.err
