#!/bin/bash
# Round-end validation on one B200: GPU parity tests (all), the paired-end bench line, launch list of the best-first bench.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 330 python -m pytest tests -m gpu -q -n 3 --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
BT_BENCH_READS=500000 timeout 150 python bench.py --policy paired --steps 3 --warmup 3 --cpu-sample 500000 > gpurun_out/bench_r1_paired.json 2> gpurun_out/bench_r1_paired.err; echo "bench paired rc=$?"; tail -c 1000 gpurun_out/bench_r1_paired.json; tail -3 gpurun_out/bench_r1_paired.err
BT_BENCH_READS=200000 BT_BENCH_NO_CPU=1 BT_BENCH_STREAMS=1 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_best.csv python bench.py --policy best --steps 2 --warmup 3 > gpurun_out/ncu_best.log 2>&1; echo "ncu rc=$?"; grep -c bt_best_kernel gpurun_out/launches_r1_best.csv
echo "total t=$(( $(date +%s) - t0 ))s"
