#!/bin/bash
O=gpurun_out/r3b; mkdir -p $O
SPH_SWEEP="lists=1,build_smem_kb=8;lists=1,build_smem_kb=10;lists=1,build_smem_kb=12;lists=1,build_smem_kb=14;lists=1,build_smem_kb=16;lists=1,build_smem_kb=20" timeout 600 python scripts/time_build.py 1e6 > $O/time_build.jsonl 2> $O/time_build.err; echo "rc=$?"; cat $O/time_build.jsonl; tail -3 $O/time_build.err
