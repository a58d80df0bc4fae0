#!/bin/bash
# round 2, call i: Δt/Δx reductions fused into the pass-2 epilogue — full GPU suite + timing
O=gpurun_out/r2i; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
SPH_SWEEP="lists=1" SPH_STEPS=120 timeout 300 python scripts/tune.py 1e6 0.15 > $O/tune.jsonl 2> $O/tune.err; echo "tune rc=$?"; cut -c1-330 $O/tune.jsonl; tail -3 $O/tune.err
