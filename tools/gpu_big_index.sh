#!/bin/bash
# One-off measurement on a larger index than travels with the repo: build a synthetic SURVEY-style genome and its
# index ON the GPU box with the reference's bowtie-build (all host cores), then run bench.py against it.
# Usage: tools/gpu_big_index.sh <Mbp> [steps]
MBP=${1:-1024}; STEPS=${2:-6}
mkdir -p gpurun_out /tmp/bigidx
python - <<PY
import sys, time
sys.path.insert(0, "tests")
import synth
from pathlib import Path
synth.CACHE = Path("/tmp/bigidx")
t = time.time()
base, _ = synth.build_synth_index("big", n_seqs=24, total_len=${MBP} * 1_000_000, seed=1, ftab_chars=10, with_gaps=True, threads=120, style="survey")
print("built", base, f"{time.time()-t:.0f}s")
open("/tmp/bigidx/base.txt", "w").write(str(base))
PY
BASE=$(cat /tmp/bigidx/base.txt)
ls -la /tmp/bigidx | head
BT_BENCH_INDEX=$BASE BT_BENCH_STREAMS=8 timeout 900 python bench.py --steps $STEPS --warmup 2 > gpurun_out/bench_big_${MBP}.json 2> gpurun_out/bench_big_${MBP}.err
tail -3 gpurun_out/bench_big_${MBP}.err
cat gpurun_out/bench_big_${MBP}.json
