#!/bin/bash
# final 2-GPU point of the scaling table
O=gpurun_out/r3f; mkdir -p $O
SPHB200_BENCH_TIMEOUT_S=600 timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 40 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -2 $O/bench_2gpu.err | cut -c1-200
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r3f/bench_2gpu.json") if l.startswith("{")][0])
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["slab"]["edges"], d["slab"]["per_rank"], d["selfcheck"]["max_rel_err_vs_single_gpu"])
PY
