#!/bin/bash
# round 2, call n: per-brick list maintenance — GPU suite (incl. on-device list verification) + timing
O=gpurun_out/r2n; mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu -x --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
SPH_SWEEP="lists=1;lists=1,list_local=0;lists=1,skin=0.15;lists=1,skin=0.2" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
