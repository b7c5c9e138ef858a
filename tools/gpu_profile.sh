#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + one full capture of the main search kernel.
mkdir -p gpurun_out
export BT_BENCH_READS=${BT_BENCH_READS:-1000000} BT_BENCH_NO_CPU=1 BT_BENCH_STREAMS=2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
BT_BENCH_STREAMS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:bt_search_kernel -s 3 -c 1 -o gpurun_out/prof_bench_r1 python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
