#!/bin/bash
# ncu --set full of the FULL list build (k_list_build + k_list_reorder right after a forced UpdateNeighbors!) at C3
O=gpurun_out/r4n; mkdir -p $O
SPH_RESET=1 SPH_PREP=0.15 timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_list_build|k_list_reorder|k_stable_rank|k_gather_table" -c 4 -f -o $O/prof_build python scripts/profile_step.py 1e6 1 > $O/prof.log 2>&1; echo "ncu rc=$?"; tail -2 $O/prof.log
