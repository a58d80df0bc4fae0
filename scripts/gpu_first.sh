#!/bin/bash
# First GPU validation round: smoke, full -m gpu suite (all failures, with measured margins),
# memcheck of one small parity test, a short bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | head -25 >> gpurun_out/host.txt; nvidia-smi >> gpurun_out/host.txt 2>&1
export SPH_PARITY_LOG=$PWD/gpurun_out/parity.jsonl
rm -f $SPH_PARITY_LOG
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "test_kernel_variants_are_bitwise_identical and c1" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -15 gpurun_out/memcheck.log
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"; cat gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
