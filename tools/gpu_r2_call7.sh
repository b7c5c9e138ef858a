#!/bin/bash
# Round 2, call 7: cold lane state in shared memory (128 / 96 registers, 4 / 5 blocks per SM) against the 160/168-register builds,
# tail block size, batch sizes; parity first.
mkdir -p gpurun_out
O=gpurun_out/c7
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_scale_parity.py -m gpu -x -q -n 3 ) > $O.pytest.log 2>&1
tail -4 $O.pytest.log
V=$PWD/bowtie_b200/variants
KB="python tools/kbench.py --index $IDX --B 2000000 --steps 6 --warmup 2 --streams 6"
$KB --tag default_mb4 --single > $O.kb.jsonl 2>$O.kb.err
for v in mb3 mb5 regs altloop; do BOWTIE_B200_LIB=$V/libbt_$v.so $KB --tag $v >> $O.kb.jsonl 2>>$O.kb.err; done
BOWTIE_B200_LIB=$V/libbt_mb5.so $KB --tag mb5_single --single >> $O.kb.jsonl 2>>$O.kb.err
for b in 2 3; do BT_MAIN_BLOCKS=$b $KB --tag mainblocks$b >> $O.kb.jsonl 2>>$O.kb.err; done
for t in 32 64; do BT_TAIL_THREADS=$t $KB --tag tailthreads$t >> $O.kb.jsonl 2>>$O.kb.err; done
BT_TAIL_THREADS=32 BT_TAIL_BLOCKS=1 $KB --tag tt32_tb1 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL_THREADS=32 BT_TAIL_BLOCKS=4 $KB --tag tt32_tb4 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL_BLOCKS=1 $KB --tag tb1 >> $O.kb.jsonl 2>>$O.kb.err
BT_TAIL_BLOCKS=3 $KB --tag tb3 >> $O.kb.jsonl 2>>$O.kb.err
for p in "8 12" "8 20" "4 12"; do set -- $p; BT_RARE_PERIOD=$1 BT_RARE_THRESH=$2 BT_HEAVY_PERIOD=$1 BT_HEAVY_THRESH=$2 $KB --tag pt$1_$2 >> $O.kb.jsonl 2>>$O.kb.err; done
python tools/kbench.py --index $IDX --B 1000000 --steps 12 --warmup 4 --streams 12 --tag B1M_s12 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 4000000 --steps 8 --warmup 2 --streams 8 --tag B4M_s8 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 8000000 --steps 6 --warmup 2 --streams 6 --tag B8M_s6 >> $O.kb.jsonl 2>>$O.kb.err
BOWTIE_B200_LIB=$V/libbt_mb5.so python tools/kbench.py --index $IDX --B 8000000 --steps 6 --warmup 2 --streams 6 --tag mb5_B8M_s6 >> $O.kb.jsonl 2>>$O.kb.err
cut -c1-330 $O.kb.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_n2k1_3g_c7.csv python tools/kbench.py --index $IDX --B 1000000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.l3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -c 2 -o gpurun_out/r2_c7_3g python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu.log 2>&1
tail -2 $O.ncu.log
