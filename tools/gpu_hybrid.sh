#!/bin/bash
# main-pass kernel / budget sweep on the bench index (development only)
run() { BT_BENCH_NO_CPU=1 BT_BENCH_STREAMS=8 timeout 300 python bench.py --steps 12 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4))"; }
BT_MAIN_KERNEL=t run "thread b24000"
BT_MAIN_KERNEL=q run "queue  b24000"
BT_MAIN_KERNEL=q BT_MAIN_BUDGET=6000 run "queue  b6000"
BT_MAIN_KERNEL=t BT_MAIN_BUDGET=6000 run "thread b6000"
BT_MAIN_KERNEL=t BT_MAIN_BUDGET=100000 run "thread b100000"
