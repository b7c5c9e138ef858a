#!/bin/bash
run() { BT_BENCH_READS=2000000 BT_BENCH_NO_CPU=1 BT_BENCH_STREAMS=8 timeout 300 python bench.py --steps 10 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), 'flags', d['config']['overflow_flags'])"; }
run "default (period 16/24, heavy 16/24, budget 8000)"
BT_HEAVY_PERIOD=4 BT_HEAVY_THRESH=8 run "heavy 4/8"
BT_HEAVY_PERIOD=1 BT_HEAVY_THRESH=1 run "heavy 1/1"
BT_HEAVY_PERIOD=64 BT_HEAVY_THRESH=30 run "heavy 64/30"
BT_MAIN_BUDGET=4000 run "budget 4000"
BT_MAIN_BUDGET=2000 run "budget 2000"
BT_RARE_PERIOD=8 BT_RARE_THRESH=20 run "main 8/20"
