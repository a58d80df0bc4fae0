#!/bin/bash
# round 2, call l: fixed sentinel records; whole GPU suite with margins; racecheck again; timing
O=gpurun_out/r2l; mkdir -p $O
rm -f $O/parity.jsonl
SPH_PARITY_LOG=$PWD/$O/parity.jsonl timeout 2400 python -m pytest tests -q -m gpu --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 $O/pytest_gpu.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_case.py 4 > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/sanitizer_racecheck.log; grep "Race reported between" $O/sanitizer_racecheck.log | sed -E 's/.*access at (void )?//; s/\(sph::Interact.*in / in /; s/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_case.py 6 > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $O/sanitizer_memcheck.log
SPH_SWEEP="lists=1" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
