#!/bin/bash
# Round 2, call 3: four-pass structure (main / tail / ultra / overflow) sweeps, register-cap variants, best-first path knobs + profile, scale parity.
mkdir -p gpurun_out
O=gpurun_out/c3
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
KB="python tools/kbench.py --index $IDX --B 2000000 --steps 6 --warmup 2 --streams 6"
$KB --tag default --single > $O.kb.jsonl 2>$O.kb.err
for v in mb4 mb5 noinner multiexit chain; do BOWTIE_B200_LIB=$PWD/bowtie_b200/variants/libbt_$v.so $KB --tag $v >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 32768 524288 0; do BT_TAIL_BUDGET=$b $KB --tag tailbudget$b >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 37 148 296; do BT_ULTRA_BLOCKS=$b $KB --tag ultra$b >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 2000 4000 16000; do BT_MAIN_BUDGET=$b $KB --tag mainbudget$b >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 1 3; do BT_TAIL_BLOCKS=$b $KB --tag tailblocks$b >> $O.kb.jsonl 2>>$O.kb.err; done
for pt in "4 8" "8 12" "12 20"; do set -- $pt; BT_RARE_PERIOD=$1 BT_RARE_THRESH=$2 BT_HEAVY_PERIOD=$1 BT_HEAVY_THRESH=$2 $KB --tag "pt$1_$2" >> $O.kb.jsonl 2>>$O.kb.err; done
python tools/kbench.py --index $IDX --B 4000000 --steps 8 --warmup 2 --streams 8 --tag B4M_s8 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 8000000 --steps 4 --warmup 1 --streams 4 --tag B8M_s4 >> $O.kb.jsonl 2>>$O.kb.err
KBB="python tools/kbench.py --index $IDX --policy best --B 1000000 --steps 4 --warmup 1 --streams 3"
for b in 12 4 8 16; do BT_BEST_BLOCKS=$b $KBB --tag best_bps$b >> $O.kb.jsonl 2>>$O.kb.err; done
KBP="python tools/kbench.py --index $IDX --policy paired --B 500000 --steps 4 --warmup 1 --streams 3"
for b in 12 4 8; do BT_BEST_BLOCKS=$b $KBP --tag paired_bps$b >> $O.kb.jsonl 2>>$O.kb.err; done
cut -c1-200 $O.kb.jsonl
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bt_best_kernel -c 1 -o gpurun_out/r2_best_3g python tools/kbench.py --index $IDX --policy best --B 500000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kbb > $O.ncu_best.log 2>&1
tail -2 $O.ncu_best.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_best_3g.csv python tools/kbench.py --index $IDX --policy best --B 500000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kbb > $O.l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_paired_3g.csv python tools/kbench.py --index $IDX --policy paired --B 250000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kbp > $O.l2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_n2k1_3g.csv python tools/kbench.py --index $IDX --B 1000000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.l3.log 2>&1
( time python -m pytest tests/test_scale_parity.py -m gpu -q -s ) > $O.scale.log 2>&1
tail -12 $O.scale.log
( time python -m pytest tests/test_gpu_parity.py tests/test_index_build.py -m gpu -x -q ) > $O.pytest.log 2>&1
tail -4 $O.pytest.log
