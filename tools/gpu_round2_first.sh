#!/bin/bash
# First GPU call of the next round: everything that was finished without a device.  Each step has its own
# timeout and log under gpurun_out/; nothing here changes clocks or sweeps memory.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
t0=$(date +%s); lap() { echo "== $1 rc=$2 t=$(( $(date +%s) - t0 ))s"; }
timeout 900 python -m pytest tests -m gpu -q -n 4 --timeout 300 > gpurun_out/r2_pytest_gpu.log 2>&1; lap "pytest -m gpu" $?; tail -3 gpurun_out/r2_pytest_gpu.log
BT_TEST_GPU_BUILD=1 timeout 600 python -m pytest tests/test_index_build.py -m gpu -q --timeout 500 > gpurun_out/r2_pytest_build.log 2>&1; lap "index builder (CUB backend)" $?; tail -3 gpurun_out/r2_pytest_build.log
timeout 300 python tools/fuzz_cli.py --gpu --iters 60 --seed 2026 > gpurun_out/r2_fuzz_gpu.log 2>&1; lap "fuzz --gpu" $?; tail -2 gpurun_out/r2_fuzz_gpu.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n2k1.json 2> gpurun_out/r2_bench_n2k1.err; lap "bench n2k1" $?; tail -c 400 gpurun_out/r2_bench_n2k1.json
timeout 200 python bench.py --policy v0 --steps 5 --warmup 3 > gpurun_out/r2_bench_v0.json 2> gpurun_out/r2_bench_v0.err; lap "bench v0 (config 2)" $?; tail -c 400 gpurun_out/r2_bench_v0.json
BT_BENCH_READS=1000000 timeout 200 python bench.py --policy best --steps 3 --warmup 3 --cpu-sample 500000 > gpurun_out/r2_bench_best.json 2> gpurun_out/r2_bench_best.err; lap "bench best" $?; tail -c 400 gpurun_out/r2_bench_best.json
BT_BENCH_READS=1000000 timeout 200 python bench.py --policy paired --steps 3 --warmup 3 --cpu-sample 300000 > gpurun_out/r2_bench_paired.json 2> gpurun_out/r2_bench_paired.err; lap "bench paired" $?; tail -c 400 gpurun_out/r2_bench_paired.json
BT_BENCH_READS=200000 BT_BENCH_STREAMS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_best.csv python bench.py --policy best --steps 2 --warmup 3 --cpu-sample 1000 > gpurun_out/r2_ncu_best.log 2>&1; lap "launch list best" $?
# A/B owed for the unified LF path (profiles/README.md): same bench with the three LF kinds on separate paths again
make -C bowtie_b200/csrc experiments > gpurun_out/r2_make_experiments.log 2>&1
BOWTIE_B200_LIB=$PWD/bowtie_b200/variants/libbt_split_lf.so timeout 200 python bench.py --steps 5 --warmup 3 --cpu-sample 1000 > gpurun_out/r2_bench_n2k1_splitlf.json 2> gpurun_out/r2_bench_n2k1_splitlf.err; lap "bench n2k1, split LF" $?; tail -c 300 gpurun_out/r2_bench_n2k1_splitlf.json
# the replay's other suggestion for the merged-LF kernel: a rare pass every 8th iteration (profiles/README.md)
BT_RARE_PERIOD=8 BT_RARE_THRESH=24 timeout 200 python bench.py --steps 5 --warmup 3 --cpu-sample 1000 > gpurun_out/r2_bench_n2k1_p8t24.json 2> gpurun_out/r2_bench_n2k1_p8t24.err; lap "bench n2k1, period 8 / threshold 24" $?; tail -c 300 gpurun_out/r2_bench_n2k1_p8t24.json
# hg19-sized index: only if the builder test above passed
if grep -q "passed" gpurun_out/r2_pytest_build.log && ! grep -q "failed" gpurun_out/r2_pytest_build.log; then
  timeout 1500 python tools/make_bench_index.py 3000 24 --gpu > gpurun_out/r2_build_3g.log 2>&1; lap "3-Gbp index" $?; tail -2 gpurun_out/r2_build_3g.log
  timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n2k1_3g.json 2> gpurun_out/r2_bench_n2k1_3g.err; lap "bench n2k1 @3G" $?; tail -c 400 gpurun_out/r2_bench_n2k1_3g.json
fi
