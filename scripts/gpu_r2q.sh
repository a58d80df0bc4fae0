#!/bin/bash
# round 2, call q: skin sweep with cheap per-brick rebuilds
O=gpurun_out/r2q; mkdir -p $O
SPH_SWEEP="lists=1,skin=0.03;lists=1,skin=0.04;lists=1,skin=0.05;lists=1,skin=0.06;lists=1,skin=0.07;lists=1,skin=0.1" SPH_STEPS=200 timeout 400 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
SPH_SWEEP="lists=1,skin=0.04;lists=1,skin=0.06;lists=1,skin=0.1" SPH_STEPS=200 timeout 400 python scripts/tune.py 1e6 0.4 > $O/tune_t04.jsonl 2> $O/tune_t04.err; echo "tune t=0.4 rc=$?"; cut -c1-330 $O/tune_t04.jsonl; tail -3 $O/tune_t04.err
