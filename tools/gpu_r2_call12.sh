#!/bin/bash
# Round 2, call 12 (2 GPUs): bench.py's multi-rank plumbing (index built once under the lock, bt_counters_allreduce over NCCL, max-over-ranks
# timing) on the shipped build, then the bowtie-compatible driver's own clock on 2 M reads.
mkdir -p gpurun_out
O=gpurun_out/c12
( time BT_BENCH_NO_CPU=1 BT_BENCH_NO_CLI=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-others ) > $O.bench2.json 2> $O.bench2.err
cut -c1-1200 $O.bench2.json; tail -5 $O.bench2.err
IDX=/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5
python - > $O.gen.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
from pathlib import Path
import bench
base = Path('/dev/shm/bowtie_b200_bench/hg19s_3000m_24_1_10_5')
g = bench.load_genome(base)
h = bench.make_reads(g, 2_000_000, seed=5)
td = Path('/dev/shm/cli12'); td.mkdir(exist_ok=True)
print(bench.write_sample(td, h, 2_000_000, False))
PY
FQ=$(ls /dev/shm/cli12/*.fq | head -1)
CLI=bowtie_b200/bowtie-b200-align
run() { local tag=$1; shift; local t0=$(date +%s.%N); env BT_CLI_TIMING=1 "$@" $CLI -n 2 -k 1 -S -x $IDX $FQ /dev/shm/cli12/out.sam 2> $O.cli_$tag.err; echo "$tag wall $(echo "$(date +%s.%N) - $t0" | bc) s"; grep -E "timing|device I/O" $O.cli_$tag.err; }
run ios4_c64
run ios1_c64 BT_CLI_IOS=1
run ios8_c32 BT_CLI_IOS=8 BT_CLI_CHUNK_MB=32
run ios4_c192 BT_CLI_CHUNK_MB=192
