#!/bin/bash
# Round 2, call 5: checkpoint slots / time-sliced tail (bt_ctxq.cuh): parity first, then A/B against the restart tail (BT_SLICES=0), slice knobs, batch sizes, launch list.
mkdir -p gpurun_out
O=gpurun_out/c5
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
( time python -m pytest tests/test_scale_parity.py -m gpu -q -s ) > $O.scale.log 2>&1
tail -9 $O.scale.log
( time python -m pytest tests/test_gpu_parity.py tests/test_best_first.py tests/test_paired.py -m gpu -x -q -n 4 ) > $O.pytest.log 2>&1
tail -4 $O.pytest.log
KB="python tools/kbench.py --index $IDX --B 2000000 --steps 6 --warmup 2 --streams 6"
$KB --tag slices_default --single > $O.kb.jsonl 2>$O.kb.err
BT_SLICES=0 $KB --tag restart_tail --single >> $O.kb.jsonl 2>>$O.kb.err
for g in 2 4 8; do BT_SLICE_GROWTH=$g $KB --tag growth$g >> $O.kb.jsonl 2>>$O.kb.err; done
for n in 3 8; do BT_SLICES=$n $KB --tag nslices$n >> $O.kb.jsonl 2>>$O.kb.err; done
for t in 32 64; do BT_SLICE_THREADS=$t $KB --tag slthreads$t >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 1 3; do BT_SLICE_BLOCKS=$b $KB --tag slblocks$b >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 2000 4000 16000; do BT_MAIN_BUDGET=$b $KB --tag mainbudget$b >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 500 4000 0; do BT_DRAIN_BUDGET=$b $KB --tag drain$b >> $O.kb.jsonl 2>>$O.kb.err; done
python tools/kbench.py --index $IDX --B 1000000 --steps 12 --warmup 4 --streams 12 --tag B1M_s12 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 2000000 --steps 12 --warmup 4 --streams 12 --tag B2M_s12 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 4000000 --steps 8 --warmup 2 --streams 8 --tag B4M_s8 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 4000000 --steps 12 --warmup 4 --streams 12 --tag B4M_s12 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 8000000 --steps 6 --warmup 2 --streams 6 --tag B8M_s6 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 16000000 --steps 4 --warmup 1 --streams 4 --tag B16M_s4 >> $O.kb.jsonl 2>>$O.kb.err
KBB="python tools/kbench.py --index $IDX --policy best --B 1000000"
$KBB --steps 4 --warmup 1 --streams 3 --tag best_B1M_s3 >> $O.kb.jsonl 2>>$O.kb.err
$KBB --steps 8 --warmup 2 --streams 6 --tag best_B1M_s6 >> $O.kb.jsonl 2>>$O.kb.err
BT_BEST_ARENA_KW=16 $KBB --steps 8 --warmup 2 --streams 6 --tag best_B1M_s6_kw16 >> $O.kb.jsonl 2>>$O.kb.err
KBP="python tools/kbench.py --index $IDX --policy paired --B 500000"
$KBP --steps 4 --warmup 1 --streams 3 --tag paired_B500k_s3 >> $O.kb.jsonl 2>>$O.kb.err
$KBP --steps 8 --warmup 2 --streams 6 --tag paired_B500k_s6 >> $O.kb.jsonl 2>>$O.kb.err
BT_BEST_T1_BLOCKS=8 $KBP --steps 8 --warmup 2 --streams 6 --tag paired_B500k_s6_t1x8 >> $O.kb.jsonl 2>>$O.kb.err
cut -c1-230 $O.kb.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_n2k1_3g_slices.csv python tools/kbench.py --index $IDX --B 1000000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.l3.log 2>&1
grep -c bt_search gpurun_out/r2_launches_n2k1_3g_slices.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -s 1 -c 1 -o gpurun_out/r2_slice0_3g python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu.log 2>&1
tail -2 $O.ncu.log
