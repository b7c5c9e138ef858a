#!/bin/bash
# Runs tools/gpu_micro.py once per tuning variant of the library (development only).
for so in bowtie_b200/variants/libbt_*.so; do
  echo "=== $so"
  BOWTIE_B200_LIB=$PWD/$so timeout 300 python tools/gpu_micro.py "$@" 2>&1 | grep '^{' 
done
