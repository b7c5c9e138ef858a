"""Child process of tests/test_scale_parity.py::test_tail_modes_agree: the library reads its tail mode once per process (BT_TAIL), so
each mode runs in its own interpreter.  Aligns n bench reads against `base` and saves found / flags / hit records to `out`."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import bowtie_b200  # noqa: E402

base, n, seed, out = Path(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
genome = bench.load_genome(base)
h = bench.make_reads(genome, n, seed=seed)
del genome
ix = bowtie_b200.Index(str(base), need_mirror=True, device=0)
found, flags, hits = ix.align(h[0], h[1], h[2], h[3], bench.lib_policy("n2k1"), slots=1, mm_cap=7)
st = ix.stats()
np.savez(out, found=found, flags=flags, hits=hits, side=np.array([st.side_fetches], np.uint64))
ix.close()
