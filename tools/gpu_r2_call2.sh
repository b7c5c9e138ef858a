#!/bin/bash
# Round 2, call 2: A/B of the search-kernel variants on the hg19-sized index, budget / deferral sweeps, ncu of the main pass, GPU test suite.
mkdir -p gpurun_out
O=gpurun_out/c2
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
KB="python tools/kbench.py --index $IDX --B 2000000 --steps 6 --warmup 2 --streams 6"
$KB --tag default --single > $O.kb.jsonl 2>$O.kb.err
for v in base sweep chase; do BOWTIE_B200_LIB=$PWD/bowtie_b200/variants/libbt_$v.so $KB --tag $v >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 2000 4000 16000; do BT_MAIN_BUDGET=$b $KB --tag budget$b >> $O.kb.jsonl 2>>$O.kb.err; done
for pt in "8 16" "4 12" "32 28" "1 1"; do set -- $pt; BT_RARE_PERIOD=$1 BT_RARE_THRESH=$2 BT_HEAVY_PERIOD=$1 BT_HEAVY_THRESH=$2 $KB --tag "pt$1_$2" >> $O.kb.jsonl 2>>$O.kb.err; done
BT_HEAVY_BLOCKS=4 $KB --tag hb4 >> $O.kb.jsonl 2>>$O.kb.err
BT_HEAVY_BLOCKS=16 $KB --tag hb16 >> $O.kb.jsonl 2>>$O.kb.err
cat $O.kb.jsonl | cut -c1-330
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -c 1 -o gpurun_out/r2_main_3g python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu.log 2>&1
tail -3 $O.ncu.log
BOWTIE_B200_LIB=$PWD/bowtie_b200/variants/libbt_base.so timeout 900 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -c 1 -o gpurun_out/r2_main_3g_base python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu2.log 2>&1
( time python -m pytest tests -m gpu -x -q --deselect tests/test_scale_parity.py ) > $O.pytest.log 2>&1
tail -5 $O.pytest.log
( time python -m pytest tests/test_scale_parity.py -m gpu -x -q ) > $O.scale.log 2>&1
tail -5 $O.scale.log
