#!/bin/bash
O=gpurun_out/r4l; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_runsimulation.py -q -m gpu > $O/pytest_runsim.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_runsim.log
