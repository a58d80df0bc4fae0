#!/bin/bash
# quick iteration: GPU parity suite + option sweep on the bench workload
mkdir -p gpurun_out
TAG=${1:-it}
export SPH_PARITY_LOG=$PWD/gpurun_out/parity_$TAG.jsonl; rm -f $SPH_PARITY_LOG
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 300 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python scripts/sweep.py > gpurun_out/sweep_$TAG.jsonl 2> gpurun_out/sweep_$TAG.err; echo "sweep rc=$?"; cat gpurun_out/sweep_$TAG.jsonl | cut -c1-260; tail -3 gpurun_out/sweep_$TAG.err
