#!/bin/bash
# round 2, call k: the whole GPU suite with parity margins logged; compute-sanitizer on the small cases
O=gpurun_out/r2k; mkdir -p $O
rm -f $O/parity.jsonl
SPH_PARITY_LOG=$PWD/$O/parity.jsonl timeout 2400 python -m pytest tests -x -q -m gpu --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -16 $O/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_case.py 6 > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_case.py 4 > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python scripts/sanitize_case.py 4 > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 $O/sanitizer_synccheck.log
