#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 26" "0 32"; do
  set -- $cfg
  SPHB200_COMPACT=$1 SPHB200_SMEM_KB=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_interact -s 6 -c 2 -f -o gpurun_out/prof_c$1_s$2 python scripts/profile_step.py 1e6 1 > gpurun_out/prof_c$1_s$2.log 2>&1; echo "ncu c$1 s$2 rc=$?"
done
