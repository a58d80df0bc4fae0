#!/bin/bash
# 2 GPUs with the final build: quick slab parity (exchange / migration / pause) + mDBC across slabs, then the bench with per-step spread
O=gpurun_out/r4g; mkdir -p $O
SLAB_PARITY_QUICK=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 scripts/slab_parity.py > $O/slab_parity_quick_2gpu.log 2>&1; echo "parity rc=$?"; grep -E "SLAB PARITY|rror" $O/slab_parity_quick_2gpu.log | tail -3
SLAB_PARITY_ONLY=mdbc timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 scripts/slab_parity.py $O/slab_parity_mdbc_2gpu.jsonl > $O/slab_parity_mdbc_2gpu.log 2>&1; echo "mdbc rc=$?"; grep -E "SLAB PARITY|rror" $O/slab_parity_mdbc_2gpu.log | tail -3
SPHB200_BENCH_TIMEOUT_S=400 timeout 450 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29603 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -2 $O/bench_2gpu.err | cut -c1-200
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r4g/bench_2gpu.json") if l.startswith("{")][0])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "b2b", d["back_to_back"]["value"], d["step_ms_spread_rank0"], d["rebuilds_in_timed_region"], d["selfcheck"]["max_rel_err_vs_single_gpu"])
PY
