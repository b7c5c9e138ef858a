"""Child process of tests/test_index_build.py::check_build_text: builds bench.py's synthetic genome with bt_index_build_text and writes
the FASTA of the same records for bowtie-build.  (A separate process so that BOWTIE_B200_LIB selects the library under test.)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import bowtie_b200  # noqa: E402

total_len, out_base, fasta = int(sys.argv[1]), sys.argv[2], sys.argv[3]
g, recs, names = bench.hg19_like_genome(total_len, 24, 7)
bowtie_b200.build_index_text(g, recs, names, out_base, off_rate=4, ftab_chars=8)
lut = np.frombuffer(b"ACGT", np.uint8)
at, k = 0, 0
with open(fasta, "wb") as f:
    for off, ln, first in recs:
        if first:
            f.write((b"" if at == 0 else b"\n") + b">" + names[k].encode() + b"\n")
            k += 1
        f.write(b"N" * off + lut[g[at:at + ln]].tobytes())
        at += ln
    f.write(b"\n")
