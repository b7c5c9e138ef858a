#!/bin/bash
# Round 2, call 8: round-robin tail (bt_tail.cu) + live-position mask: parity, A/B against the restart tail and the previous build, knobs, bench line.
mkdir -p gpurun_out
O=gpurun_out/c8
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_scale_parity.py tests/test_device_io.py tests/test_cli_parity.py -m gpu -x -q -n 3 ) > $O.pytest.log 2>&1
tail -4 $O.pytest.log
V=$PWD/bowtie_b200/variants
KB="python tools/kbench.py --index $IDX --B 2000000 --steps 6 --warmup 2 --streams 6"
$KB --tag rr_default --single > $O.kb.jsonl 2>$O.kb.err
BT_TAIL=restart $KB --tag restart --single >> $O.kb.jsonl 2>>$O.kb.err
BOWTIE_B200_LIB=$V/libbt_base.so $KB --tag base_nomask_restart >> $O.kb.jsonl 2>>$O.kb.err
$KB --tag rr_default_again >> $O.kb.jsonl 2>>$O.kb.err
for q in 1024 16384 65536; do BT_TAIL_QUANTUM=$q $KB --tag quantum$q >> $O.kb.jsonl 2>>$O.kb.err; done
for w in 4 8; do BT_TAIL_WARPS=$w $KB --tag tailwarps$w >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 2000 4000 16000; do BT_MAIN_BUDGET=$b $KB --tag mainbudget$b >> $O.kb.jsonl 2>>$O.kb.err; done
BT_MAIN_BUDGET=4000 BT_DRAIN_BUDGET=500 $KB --tag mb4000_drain500 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 1000000 --steps 12 --warmup 4 --streams 12 --tag B1M_s12 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 1000000 --steps 12 --warmup 4 --streams 4 --tag B1M_s4 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 4000000 --steps 8 --warmup 2 --streams 8 --tag B4M_s8 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 8000000 --steps 6 --warmup 2 --streams 6 --tag B8M_s6 >> $O.kb.jsonl 2>>$O.kb.err
cut -c1-330 $O.kb.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_n2k1_3g_c8.csv python tools/kbench.py --index $IDX --B 1000000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.l3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"bt_search_kernel|bt_tail_kernel" -c 2 -o gpurun_out/r2_c8_3g python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu.log 2>&1
tail -2 $O.ncu.log
( time timeout 1200 python bench.py ) > $O.bench.json 2> $O.bench.err
cut -c1-400 $O.bench.json
