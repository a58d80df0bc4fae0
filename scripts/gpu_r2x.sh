#!/bin/bash
# round 2, call x: epilogue reads prefetched at sub-brick start
O=gpurun_out/r2x; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_lists.py tests/test_gpu_parity.py -q -m gpu -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
SPH_SWEEP="lists=1;lists=1" SPH_STEPS=200 timeout 400 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
