#!/bin/bash
# bricks halved only until there are two per SM: full GPU suite (no -x), small configs, C3 bench
O=gpurun_out/r4j; mkdir -p $O
rm -f $O/parity.jsonl
SPH_PARITY_LOG=$PWD/$O/parity.jsonl timeout 700 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 200 python scripts/small_profile.py > $O/small_profile.jsonl 2> $O/small_profile.err; echo "profile rc=$?"; cut -c1-200 $O/small_profile.jsonl; tail -3 $O/small_profile.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
