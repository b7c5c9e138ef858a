#!/bin/bash
# Round 2, call 11: where the bowtie-compatible driver's time goes on the device I/O path (chunks in flight, chunk size); the tail-mode
# agreement test; a 2-rank pass of bench.py's multi-GPU plumbing is a separate call.
mkdir -p gpurun_out
O=gpurun_out/c11
python -c "import bench; print(bench.ensure_index(3000, 0))" > $O.index.log 2>&1
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
python - > $O.gen.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
from pathlib import Path
import bench
base = Path('/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5')
g = bench.load_genome(base)
t = time.time(); h = bench.make_reads(g, 2_000_000, seed=5); print('make_reads 2M', round(time.time() - t, 1), 's')
td = Path('/dev/shm/cli11'); td.mkdir(exist_ok=True)
print(bench.write_sample(td, h, 2_000_000, False))
PY
cat $O.gen.log
FQ=$(ls /dev/shm/cli11/*.fq | head -1)
CLI=bowtie_b200/bowtie-b200-align
run() { local tag=$1; shift; /usr/bin/time -f "$tag wall %e s" env BT_CLI_TIMING=1 "$@" $CLI -n 2 -k 1 -S -x $IDX $FQ /dev/shm/cli11/out.sam 2> $O.cli_$tag.err; grep -E "wall|device I/O|host pipeline" $O.cli_$tag.err; }
head -c 1000 $FQ > /dev/shm/cli11/one.fq
/usr/bin/time -f "one-read wall %e s" $CLI -n 2 -k 1 -S -x $IDX /dev/shm/cli11/one.fq /dev/shm/cli11/one.sam 2>&1 | tail -1
run ios4_c64
run ios1_c64 BT_CLI_IOS=1
run ios2_c64 BT_CLI_IOS=2
run ios8_c64 BT_CLI_IOS=8
run ios4_c32 BT_CLI_CHUNK_MB=32
run ios4_c192 BT_CLI_CHUNK_MB=192
run ios8_c32 BT_CLI_IOS=8 BT_CLI_CHUNK_MB=32
run hostio BT_CLI_HOST_IO=1
( time timeout 600 python -m pytest tests/test_scale_parity.py -m gpu -q -k tail_modes ) > $O.pytest.log 2>&1
tail -3 $O.pytest.log
