#!/bin/bash
# round 2, call j (2 GPUs): slab parity with the ring kernel + fused reductions; bench.py contract at N=1 and N=2
O=gpurun_out/r2j; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/host.txt; nproc >> $O/host.txt
timeout 900 python -m pytest tests/test_gpu_slab.py -x -q -m gpu > $O/pytest_slab.log 2>&1; echo "pytest slab rc=$?"; tail -5 $O/pytest_slab.log
timeout 900 python bench.py --steps 40 --warmup 5 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench1 rc=$?"; tail -3 $O/bench_1gpu.err; python -c "
import json; d=json.load(open('$O/bench_1gpu.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['stage_ms']); print(d['roofline']['frac'], d['roofline_fp32']['frac'], d['e2e']['value'], d['cpu_baseline'])"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 40 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; tail -5 $O/bench_2gpu.err; python -c "
import json; d=json.load(open('$O/bench_2gpu.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['stage_ms']); print(d['slab']); print(d['selfcheck'])"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -2 $O/bench_ref.err; cut -c1-600 $O/bench_ref.json
