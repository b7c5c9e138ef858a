#!/bin/bash
# Round 2, call 6 (re-run of call 5 after the container was replaced): checkpoint slots / time-sliced tail on the device.
# Parity first, then A/B against the restart tail (BT_SLICES=0), slice knobs, batch sizes, launch list, ncu of main pass and slice 0, bench line.
mkdir -p gpurun_out
O=gpurun_out/c6
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
( time timeout 900 python -m pytest tests/test_scale_parity.py -m gpu -q -s ) > $O.scale.log 2>&1
tail -6 $O.scale.log
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_best_first.py tests/test_paired.py tests/test_device_io.py tests/test_index_build.py -m gpu -x -q -n 4 ) > $O.pytest.log 2>&1
tail -4 $O.pytest.log
KB="python tools/kbench.py --index $IDX --B 2000000 --steps 6 --warmup 2 --streams 6"
$KB --tag slices_default --single > $O.kb.jsonl 2>$O.kb.err
BT_SLICES=0 $KB --tag restart_tail --single >> $O.kb.jsonl 2>>$O.kb.err
for g in 2 4; do BT_SLICE_GROWTH=$g $KB --tag growth$g >> $O.kb.jsonl 2>>$O.kb.err; done
for n in 3 8; do BT_SLICES=$n $KB --tag nslices$n >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 1 3; do BT_SLICE_BLOCKS=$b $KB --tag slblocks$b >> $O.kb.jsonl 2>>$O.kb.err; done
python tools/kbench.py --index $IDX --B 1000000 --steps 12 --warmup 4 --streams 12 --tag B1M_s12 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 4000000 --steps 8 --warmup 2 --streams 8 --tag B4M_s8 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --policy best --B 1000000 --steps 8 --warmup 2 --streams 6 --tag best_B1M_s6 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --policy paired --B 500000 --steps 8 --warmup 2 --streams 6 --tag paired_B500k_s6 >> $O.kb.jsonl 2>>$O.kb.err
cut -c1-260 $O.kb.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_n2k1_3g_slices.csv python tools/kbench.py --index $IDX --B 1000000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.l3.log 2>&1
grep -c bt_search gpurun_out/r2_launches_n2k1_3g_slices.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -c 3 -o gpurun_out/r2_slices_3g python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu.log 2>&1
tail -2 $O.ncu.log
( time timeout 1200 python bench.py ) > $O.bench.json 2> $O.bench.err
cut -c1-600 $O.bench.json
