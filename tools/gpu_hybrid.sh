#!/bin/bash
run() { BT_BENCH_NO_CPU=1 BT_BENCH_STREAMS=8 timeout 300 python bench.py --steps 12 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4), 'flags', d['config']['overflow_flags'])"; }
run "q/q default"
BT_HEAVY_BUDGET=50000 run "q/q hb50000"
BT_MAIN_BUDGET=4000 run "q/q mb4000"
BT_HEAVY_KERNEL=t run "q/t"
BT_MAIN_KERNEL=t BT_HEAVY_KERNEL=t run "t/t"
