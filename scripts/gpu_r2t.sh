#!/bin/bash
# round 2, call t: lookahead batching of per-brick list rebuilds
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_lists.py tests/test_gpu_zz_fullsize.py -q -m gpu -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
SPH_SWEEP="lists=1,list_lookahead=0;lists=1,list_lookahead=2;lists=1,list_lookahead=3;lists=1,list_lookahead=5;lists=1,list_lookahead=8;lists=1,list_lookahead=5,skin=0.05;lists=1,list_lookahead=8,skin=0.03" SPH_STEPS=200 timeout 500 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
