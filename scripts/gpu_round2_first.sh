#!/bin/bash
# First GPU visit of round 2 (ONE GPU; every command under its own short timeout — a hung multi-GPU
# experiment burnt 100 GPU-minutes in round 1).  Validates the final round-1 commit, then measures
# the experimental paths that could not be run on hardware in round 1.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh r2a'
TAG=${1:-r2a}
O=gpurun_out/$TAG; mkdir -p $O
export SPH_PARITY_LOG=$PWD/$O/parity.jsonl; rm -f $SPH_PARITY_LOG
echo "== pytest gpu (incl. the full-size and experimental-path tests, which round 1 could not run)"
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 200 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu.log | cut -c1-300
echo "== list entry order: window order vs bank-aware (sph_listorder.h)"
SPH_SWEEP="lists=1;lists=1,list_order=1" timeout 120 python scripts/list_diag.py > $O/list_order.jsonl 2> $O/list_order.err; cut -c1-330 $O/list_order.jsonl; tail -2 $O/list_order.err
SPH_VEL=2 SPH_SWEEP="lists=1;lists=1,list_order=1;lists=1,list_order=1,skin=0.07" timeout 120 python scripts/list_diag.py > $O/list_order_vel2.jsonl 2>> $O/list_order.err; cut -c1-330 $O/list_order_vel2.jsonl
echo "== bench"
timeout 200 python bench.py --steps 100 --warmup 10 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
ls -la $O
# then, separately and with --gpus 2 --timeout 150:
#   SPHB200_SLAB_WAITVALUE=1 SLAB_PARITY_QUICK=1 timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#       --master-addr 127.0.0.1 --master-port 29511 scripts/slab_parity.py
