#!/bin/bash
# mDBC in slab mode: world-of-one tests + the single-GPU mDBC test (refactored gather), then C5 on 2 GPUs
O=gpurun_out/r4a; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py -q -m gpu -k "mdbc" -x > $O/pytest_mdbc.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_mdbc.log
SLAB_PARITY_ONLY=mdbc timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 scripts/slab_parity.py $O/slab_parity_mdbc_2gpu.jsonl > $O/slab_parity_mdbc_2gpu.log 2>&1; echo "parity rc=$?"; grep -E "^\{|SLAB PARITY|rror" $O/slab_parity_mdbc_2gpu.log | cut -c1-900 | tail -8
