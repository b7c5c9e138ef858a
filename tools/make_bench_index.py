#!/usr/bin/env python
"""Builds the benchmark index with the reference's own bowtie-build (oracle/_ref/bowtie-build-s):
a synthetic genome in the style of SURVEY.md §8(d) (repeat families, 1-15 % divergence, N gaps).
Usage: python tools/make_bench_index.py [Mbp=128] [n_seqs=24] [--gpu]
--gpu: build with bt_index_build (the suffix sort on the device) instead — the way to an hg19-sized (3000 Mbp) index on a GPU box,
where nothing persists between calls and the reference's builder would take hours."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from synth import build_synth_index  # noqa: E402

gpu = "--gpu" in sys.argv
args = [a for a in sys.argv[1:] if a != "--gpu"]
mbp = int(args[0]) if len(args) > 0 else 128
nseq = int(args[1]) if len(args) > 1 else 24
sys.path.insert(0, str(ROOT))
t = time.time()
base, _ = build_synth_index("benchs", n_seqs=nseq, total_len=mbp * 1_000_000, seed=1, ftab_chars=10, with_gaps=True, style="survey", builder="gpu" if gpu else "reference")
print(base, f"{time.time() - t:.0f}s")
