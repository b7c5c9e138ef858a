#!/usr/bin/env python
"""Builds the benchmark index with the reference's own bowtie-build (oracle/_ref/bowtie-build-s):
a synthetic genome in the style of SURVEY.md §8(d) (repeat families, 1-15 % divergence, N gaps).
Usage: python tools/make_bench_index.py [Mbp=128] [n_seqs=24]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
from synth import build_synth_index  # noqa: E402

mbp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 24
t = time.time()
base, _ = build_synth_index("benchs", n_seqs=nseq, total_len=mbp * 1_000_000, seed=1, ftab_chars=10, with_gaps=True, style="survey")
print(base, f"{time.time() - t:.0f}s")
