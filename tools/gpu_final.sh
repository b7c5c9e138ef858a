#!/bin/bash
# Last GPU call of the round (short budget): best-first bench after the lazy range-state blocks, the paired-end / best-first GPU tests
# (PairedBWAlignerV2, four arena tiers), the paired bench.
mkdir -p gpurun_out
t0=$(date +%s)
BT_BENCH_READS=500000 timeout 75 python bench.py --policy best --steps 3 --warmup 3 --cpu-sample 500000 > gpurun_out/bench_r1_best.json 2> gpurun_out/bench_r1_best.err; echo "bench best rc=$? t=$(( $(date +%s) - t0 ))s"; tail -c 600 gpurun_out/bench_r1_best.json
timeout 130 python -m pytest tests/test_paired.py tests/test_best_first.py -m gpu -q -n 4 --timeout 100 > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/pytest_gpu2.log
BT_BENCH_READS=500000 timeout 70 python bench.py --policy paired --steps 3 --warmup 3 --cpu-sample 300000 > gpurun_out/bench_r1_paired.json 2> gpurun_out/bench_r1_paired.err; echo "bench paired rc=$? t=$(( $(date +%s) - t0 ))s"; tail -c 500 gpurun_out/bench_r1_paired.json
