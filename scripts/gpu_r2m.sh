#!/bin/bash
# round 2, call m (8 GPUs): C4 (16.3 M particles, 8 y-slabs) through bench.py; reference arm code path at N = 8 (short budget)
O=gpurun_out/r3e; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/host.txt; nproc >> $O/host.txt
SPHB200_BENCH_TIMEOUT_S=900 timeout 1000 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 40 --warmup 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?"; tail -4 $O/bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r3e/bench_8gpu.json") if l.startswith("{")][0])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")})
for k, v in d["stage_ms"].items(): print(f"  {v:8.4f}  {k}")
print(d["slab"]); print(d["selfcheck"]); print(d["back_to_back"]); print("e2e", d["e2e"]["value"]); print(d["clocks"])
PY
