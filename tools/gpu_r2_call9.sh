#!/bin/bash
# Round 2, call 9: round-robin tail with adaptive packing (BT_TAIL=rr, wtarget / mincap) against the restart tail, longer runs and repeats.
mkdir -p gpurun_out
O=gpurun_out/c9
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
( time BT_TAIL=rr timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_scale_parity.py -m gpu -x -q -n 3 ) > $O.pytest_rr.log 2>&1
tail -3 $O.pytest_rr.log
KB="timeout 300 python tools/kbench.py --index $IDX --B 2000000 --steps 12 --warmup 2 --streams 6"
$KB --tag restart_1 --single > $O.kb.jsonl 2>$O.kb.err
BT_TAIL=rr $KB --tag rr_w256_c4_1 --single >> $O.kb.jsonl 2>>$O.kb.err
$KB --tag restart_2 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr $KB --tag rr_w256_c4_2 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr BT_TAIL_WTARGET=512 $KB --tag rr_w512_c4 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr BT_TAIL_WTARGET=128 $KB --tag rr_w128_c4 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr BT_TAIL_MINCAP=1 $KB --tag rr_w256_c1 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr BT_TAIL_MINCAP=8 $KB --tag rr_w256_c8 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr BT_TAIL_WTARGET=1024 BT_TAIL_MINCAP=1 $KB --tag rr_w1024_c1 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL_THREADS=32 $KB --tag restart_tt32 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL_THREADS=32 BT_TAIL_BLOCKS=1 $KB --tag restart_tt32_tb1 >> $O.kb.jsonl 2>>$O.kb.err
KB8="timeout 300 python tools/kbench.py --index $IDX --B 8000000 --steps 6 --warmup 2 --streams 6"
$KB8 --tag restart_B8M >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr $KB8 --tag rr_B8M >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL=rr BT_TAIL_MINCAP=1 BT_TAIL_WTARGET=1024 $KB8 --tag rr_w1024_c1_B8M >> $O.kb.jsonl 2>>$O.kb.err
cut -c1-330 $O.kb.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_n2k1_3g_c9_rr.csv env BT_TAIL=rr python tools/kbench.py --index $IDX --B 1000000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.l3.log 2>&1
