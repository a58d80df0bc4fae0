#!/bin/bash
# round 2, call w (2 GPUs): general slab axis (x-slabs) — single-GPU suite + 2-GPU slab parity on every axis
O=gpurun_out/r2w; mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 scripts/slab_parity.py $O/slab_parity_2gpu.jsonl > $O/slab_parity.log 2>&1; echo "slab parity rc=$?"; tail -3 $O/slab_parity.log | cut -c1-200
python - <<'PY'
import json
for l in open("gpurun_out/r2w/slab_parity_2gpu.jsonl"):
    d = json.loads(l)
    if "skipped" in d: print("skipped", d); continue
    print(d["ok"], d["case"], d["float"], "axis", d["axis"], "vs single", max(d[f"err_{f}_vs_single"] for f in ("Position","Velocity","Density")), "vs oracle", max(d[f"err_{f}_vs_oracle"] for f in ("Position","Velocity","Density")))
PY
