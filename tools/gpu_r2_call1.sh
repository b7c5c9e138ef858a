#!/bin/bash
# Round 2, first GPU call: first device run of the index builder, the hg19-sized index, the bench on it, ncu of the search kernel.
mkdir -p gpurun_out
export BT_BUILD_VERBOSE=1
( time python -m pytest tests/test_index_build.py -m gpu -x -q ) > gpurun_out/c1_build_test.log 2>&1
tail -3 gpurun_out/c1_build_test.log
( time python bench.py --make-index /dev/shm/bowtie_b200_bench/probe --mbp 256 ) > gpurun_out/c1_build256.log 2>&1
tail -30 gpurun_out/c1_build256.log
rm -f /dev/shm/bowtie_b200_bench/probe.*
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 3000 gpurun_out/c1_bench.err; cat gpurun_out/c1_bench.json
ls -la /dev/shm/bowtie_b200_bench/ > gpurun_out/c1_ls.log; cat /dev/shm/bowtie_b200_bench/*.json
export BT_BENCH_NO_CPU=1 BT_BENCH_NO_STRONG=1 BT_BENCH_NO_OTHERS=1
BT_BENCH_READS=1000000 BT_BENCH_STREAMS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -s 2 -c 1 -o gpurun_out/r2_search_3g python bench.py --steps 1 --warmup 1 > gpurun_out/c1_ncu.log 2>&1
tail -5 gpurun_out/c1_ncu.log
BT_BENCH_READS=1000000 BT_BENCH_STREAMS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_3g.csv python bench.py --steps 2 --warmup 1 > gpurun_out/c1_launches.log 2>&1
tail -3 gpurun_out/c1_launches.log
