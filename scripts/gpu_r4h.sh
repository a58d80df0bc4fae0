#!/bin/bash
O=gpurun_out/r4h; mkdir -p $O
timeout 200 python scripts/e2e_anatomy.py > $O/e2e_anatomy.json 2> $O/e2e_anatomy.err; echo "anatomy rc=$?"; cat $O/e2e_anatomy.json; tail -2 $O/e2e_anatomy.err
SMALL_PROFILE_NCU=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_interact_ring" -s 60 -c 2 -f -o $O/prof_ring_c1 python scripts/small_profile.py c1 > $O/ncu_c1.log 2>&1; echo "ncu rc=$?"
