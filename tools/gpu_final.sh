#!/bin/bash
# Round-end validation on one B200: GPU parity tests, smoke, the headline bench line and the best-first bench line.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q -n 3 --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 420 python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; echo "bench rc=$?"; tail -c 700 gpurun_out/bench_r1_n1.json
BT_BENCH_READS=1000000 timeout 300 python bench.py --policy best --steps 3 --warmup 3 > gpurun_out/bench_r1_best.json 2> gpurun_out/bench_r1_best.err; echo "bench best rc=$?"; tail -c 900 gpurun_out/bench_r1_best.json; tail -3 gpurun_out/bench_r1_best.err
echo "total t=$(( $(date +%s) - t0 ))s"
