#!/bin/bash
# lean sequence without standby cull kernels, S19 merged into the next head: tests, small configs
O=gpurun_out/r4f; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_lean.py tests/test_gpu_lists.py tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_runsimulation.py -q -m gpu -x > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest.log
timeout 200 python scripts/small_profile.py > $O/small_profile.jsonl 2> $O/small_profile.err; echo "profile rc=$?"; cut -c1-200 $O/small_profile.jsonl; tail -3 $O/small_profile.err
