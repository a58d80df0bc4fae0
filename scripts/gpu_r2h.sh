#!/bin/bash
# round 2, call h: NCW / LIST_PF variants of the packed ring kernel
O=gpurun_out/r2h; mkdir -p $O
for v in ncw12 ncw10 pf4 pf4ncw12 pf2; do
  SPHB200_LIB=sphexample_b200/lib/libsphb200_$v.so SPH_SWEEP="lists=1" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune_$v.jsonl 2> $O/tune_$v.err; echo "tune $v rc=$?"; cut -c1-200 $O/tune_$v.jsonl; tail -2 $O/tune_$v.err
done
SPH_SWEEP="lists=1" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune base rc=$?"; cut -c1-200 $O/tune.jsonl
