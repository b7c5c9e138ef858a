#!/bin/bash
# Round 2, call 13 (last GPU minutes): the device I/O tests on the final build and the driver's own clock on 2 M reads.
mkdir -p gpurun_out
O=gpurun_out/c13
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
python - > $O.gen.log 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from pathlib import Path
import bench
base = Path('/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5')
g = bench.load_genome(base)
h = bench.make_reads(g, 2_000_000, seed=5)
td = Path('/dev/shm/cli13'); td.mkdir(exist_ok=True)
print(bench.write_sample(td, h, 2_000_000, False))
PY
FQ=$(ls /dev/shm/cli13/*.fq | head -1)
CLI=bowtie_b200/bowtie-b200-align
for tag in a b; do BT_CLI_TIMING=1 timeout 60 $CLI -n 2 -k 1 -S -x $IDX $FQ /dev/shm/cli13/out.sam 2> $O.cli_$tag.err; echo "run $tag rc=$?"; grep -E "timing|device I/O" $O.cli_$tag.err; done
BT_CLI_TIMING=1 timeout 60 oracle/_ref/bowtie-align-s -n 2 -k 1 -S -p 16 -t -x $IDX $FQ /dev/shm/cli13/ref.sam 2> $O.ref.err; grep -E "Time|reads" $O.ref.err | head -5
grep -v "^@PG" /dev/shm/cli13/out.sam | sort | md5sum; grep -v "^@PG" /dev/shm/cli13/ref.sam | sort | md5sum
( timeout 100 python -m pytest tests/test_device_io.py -m gpu -q -x ) > $O.pytest.log 2>&1; tail -2 $O.pytest.log
