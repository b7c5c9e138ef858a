#!/bin/bash
# Round 2, call 10: the whole -m gpu suite, smoke(), the differential fuzzer on the device, launch list + full ncu captures of the shipped
# main and tail passes, batch sizes of the other two engines, the bench line.
mkdir -p gpurun_out
O=gpurun_out/c10
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
( time timeout 1500 python -m pytest tests -m gpu -q -n 4 ) > $O.pytest.log 2>&1
tail -4 $O.pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O.smoke.log 2>&1
tail -3 $O.smoke.log
( time timeout 400 python tools/fuzz_cli.py --gpu --iters 60 --seed 777 ) > $O.fuzz.log 2>&1
tail -2 $O.fuzz.log
timeout 300 python tools/kbench.py --index $IDX --policy best --B 2000000 --steps 6 --warmup 2 --streams 6 --tag best_B2M_s6 > $O.kb.jsonl 2>$O.kb.err
timeout 300 python tools/kbench.py --index $IDX --policy best --B 1000000 --steps 8 --warmup 2 --streams 8 --tag best_B1M_s8 >> $O.kb.jsonl 2>>$O.kb.err
timeout 300 python tools/kbench.py --index $IDX --policy paired --B 1000000 --steps 6 --warmup 2 --streams 6 --tag paired_B1M_s6 >> $O.kb.jsonl 2>>$O.kb.err
timeout 300 python tools/kbench.py --index $IDX --B 16000000 --steps 4 --warmup 1 --streams 4 --tag n2k1_B16M_s4 >> $O.kb.jsonl 2>>$O.kb.err
cut -c1-300 $O.kb.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches_n2k1_3g_final.csv python tools/kbench.py --index $IDX --B 1000000 --steps 2 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.l3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -c 2 -o gpurun_out/r2_final_3g python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu.log 2>&1
tail -2 $O.ncu.log
( time timeout 1500 python bench.py ) > $O.bench.json 2> $O.bench.err
cut -c1-500 $O.bench.json; tail -4 $O.bench.err
