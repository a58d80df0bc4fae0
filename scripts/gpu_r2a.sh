#!/bin/bash
# round 2, call a: baseline sweep of the existing options in developed flow (incl. bank-aware list order)
O=gpurun_out/r2a; mkdir -p $O
(nvidia-smi; nproc; lscpu | head -20) > $O/host.txt 2>&1
SPH_SWEEP="lists=1;lists=1,list_order=1;lists=1,list_order=1,skin=0.07;lists=1,list_order=1,skin=0.15;lists=1,list_order=1,list_smem_kb=56;lists=1,list_order=1,list_smem_kb=100" SPH_STEPS=120 timeout 600 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cat $O/tune.jsonl; tail -3 $O/tune.err
SPHB200_LIST_ORDER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_interact_list" -s 8 -c 2 -f -o $O/prof_interact_lo1 python scripts/profile_step.py 1e6 2 > $O/prof.log 2>&1; echo "ncu rc=$?"; tail -2 $O/prof.log
