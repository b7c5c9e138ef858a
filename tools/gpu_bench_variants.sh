#!/bin/bash
# micro + bench once per tuning variant (development only)
for so in bowtie_b200/variants/libbt_*.so; do
  echo "=== $so"
  BOWTIE_B200_LIB=$PWD/$so timeout 300 python tools/gpu_micro.py ecoli 2>&1 | grep '^{' | cut -c1-200
  BOWTIE_B200_LIB=$PWD/$so BT_BENCH_NO_CPU=1 BT_BENCH_STREAMS=8 timeout 300 python bench.py --steps 10 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4))"
done
