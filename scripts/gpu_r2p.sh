#!/bin/bash
# round 2, call p: warp-per-brick bounds kernel; skin sweep under per-brick list maintenance
O=gpurun_out/r2p; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_lists.py -q -m gpu -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
SPH_SWEEP="lists=1;lists=1,skin=0.07;lists=1,skin=0.13;lists=1,list_local=0" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
