#!/bin/bash
# round 2, call u (4 GPUs): per-rank pass times for the slab load balance, two boundary weights
O=gpurun_out/r2u; mkdir -p $O
for w in 0.5 1.0; do
SPHB200_BOUNDARY_WEIGHT=$w SPHB200_BENCH_TIMEOUT_S=600 timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 40 --warmup 5 --no-selfcheck > $O/bench_4gpu_w$w.json 2> $O/bench_4gpu_w$w.err; echo "bench4 w=$w rc=$?"; tail -2 $O/bench_4gpu_w$w.err | cut -c1-200
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2u/bench_4gpu_w$w.json") if l.startswith("{")][0])
print({k: d[k] for k in ("value", "ms_per_step")}, d["slab"]["edges"], d["slab"]["per_rank"])
PY
done
