#!/bin/bash
# Round 2, call 4: search-kernel A/B with the tail pass unbudgeted (the configuration that won call 3), batch-size effects, best-first knobs,
# ncu of the main pass (shipped vs variants), device-I/O tests, CLI end to end.
mkdir -p gpurun_out
O=gpurun_out/c4
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
KB="python tools/kbench.py --index $IDX --B 2000000 --steps 6 --warmup 2 --streams 6"
$KB --tag default --single > $O.kb.jsonl 2>$O.kb.err
$KB --tag default_again >> $O.kb.jsonl 2>>$O.kb.err
for v in noinner multiexit chain splitchase r1like; do BOWTIE_B200_LIB=$PWD/bowtie_b200/variants/libbt_$v.so $KB --tag $v >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 2000 4000 16000 32000; do BT_MAIN_BUDGET=$b $KB --tag mainbudget$b >> $O.kb.jsonl 2>>$O.kb.err; done
for b in 1 3; do BT_TAIL_BLOCKS=$b $KB --tag tailblocks$b >> $O.kb.jsonl 2>>$O.kb.err; done
for pt in "4 8" "16 24" "2 4"; do set -- $pt; BT_RARE_PERIOD=$1 BT_RARE_THRESH=$2 BT_HEAVY_PERIOD=$1 BT_HEAVY_THRESH=$2 $KB --tag "pt$1_$2" >> $O.kb.jsonl 2>>$O.kb.err; done
for pt in "4 8" "1 1"; do set -- $pt; BT_HEAVY_PERIOD=$1 BT_HEAVY_THRESH=$2 $KB --tag "heavy_pt$1_$2" >> $O.kb.jsonl 2>>$O.kb.err; done
python tools/kbench.py --index $IDX --B 4000000 --steps 8 --warmup 2 --streams 8 --tag B4M_s8 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 8000000 --steps 6 --warmup 2 --streams 6 --tag B8M_s6 >> $O.kb.jsonl 2>>$O.kb.err
python tools/kbench.py --index $IDX --B 16000000 --steps 4 --warmup 1 --streams 4 --tag B16M_s4 >> $O.kb.jsonl 2>>$O.kb.err
KBB="python tools/kbench.py --index $IDX --policy best --steps 4 --warmup 1 --streams 3"
$KBB --B 1000000 --tag best_B1M >> $O.kb.jsonl 2>>$O.kb.err
$KBB --B 2000000 --tag best_B2M >> $O.kb.jsonl 2>>$O.kb.err
BT_BEST_ARENA_KW=32 $KBB --B 1000000 --tag best_B1M_kw32 >> $O.kb.jsonl 2>>$O.kb.err
BT_BEST_T1_BLOCKS=4 $KBB --B 1000000 --tag best_B1M_t1x4 >> $O.kb.jsonl 2>>$O.kb.err
KBP="python tools/kbench.py --index $IDX --policy paired --steps 4 --warmup 1 --streams 3"
$KBP --B 500000 --tag paired_B500k >> $O.kb.jsonl 2>>$O.kb.err
BT_BEST_ARENA_KW=32 $KBP --B 500000 --tag paired_kw32 >> $O.kb.jsonl 2>>$O.kb.err
BT_BEST_T1_BLOCKS=4 $KBP --B 500000 --tag paired_t1x4 >> $O.kb.jsonl 2>>$O.kb.err
BT_BEST_ARENA_KW=32 BT_BEST_T1_BLOCKS=4 $KBP --B 1000000 --tag paired_B1M_kw32_t1x4 >> $O.kb.jsonl 2>>$O.kb.err
cut -c1-160 $O.kb.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -c 1 -o gpurun_out/r2_main_3g_v2 python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu.log 2>&1
BOWTIE_B200_LIB=$PWD/bowtie_b200/variants/libbt_noinner.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -c 1 -o gpurun_out/r2_main_3g_noinner python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bt_search_kernel -s 1 -c 1 -o gpurun_out/r2_tail_3g_v2 python tools/kbench.py --index $IDX --B 1000000 --steps 1 --warmup 0 --streams 1 --reads /dev/shm/kb1m > $O.ncu3.log 2>&1
( time python -m pytest tests/test_device_io.py tests/test_cli_parity.py -m gpu -x -q ) > $O.pytest.log 2>&1
tail -4 $O.pytest.log
python - <<'PY' > $O.cli.log 2>&1
import bench, tempfile, time, subprocess, json
from pathlib import Path
base = Path("/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5")
g = bench.load_genome(base)
h = bench.make_reads(g, 2_000_000, seed=5)
del g
with tempfile.TemporaryDirectory(dir="/dev/shm") as tdn:
    td = Path(tdn)
    rr = bench.ReferenceRunner(base, ["-n", "2", "-k", "1"], False, td, h)
    rr.pick_threads(200_000)
    print(json.dumps(bench.cli_e2e(base, h, 2_000_000, td, rr)))
    fq = td / "cli" / "s.fq"
    for env in ({"BT_CLI_TIMING": "1"}, {"BT_CLI_TIMING": "1", "BT_CLI_HOST_IO": "1"}):
        import os
        t0 = time.time(); p = subprocess.run([str(bench.ROOT / "bowtie_b200" / "bowtie-b200-align"), "-n", "2", "-k", "1", "-S", "-x", str(base), str(fq), str(td / "x.sam")], capture_output=True, text=True, env=dict(os.environ, **env))
        print(env, round(time.time() - t0, 2), p.stderr[-600:])
PY
tail -12 $O.cli.log | cut -c1-400
