#!/bin/bash
# call r3a: cost of a full list build (first step after an upload) vs build / reorder shared-memory settings
O=gpurun_out/r3a; mkdir -p $O
SPH_SWEEP="lists=1;lists=1,reorder_slots=192;lists=1,reorder_slots=288;lists=1,build_smem_kb=36;lists=1,build_smem_kb=48;lists=1,build_smem_kb=16;lists=1,list_reorder=0" timeout 600 python scripts/time_build.py 1e6 > $O/time_build.jsonl 2> $O/time_build.err; echo "rc=$?"; cat $O/time_build.jsonl; tail -3 $O/time_build.err
