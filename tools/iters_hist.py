"""Distribution of the search cost per read (transitions of the lane state machine) on the bench workload, from the host emulation of
the device code (tests/host_emu): quantiles, how many reads exceed a budget, and how the total splits between the main pass and the
tail for a given main budget.  Development aid for the tail's design; not a measurement.
Usage: BT_BENCH_INDEX=<base> python tools/iters_hist.py [n_reads=200000] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from helpers import HostEmu, Policy  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
base, name = bench.fallback_index()
genome = bench.load_genome(base)
emu = HostEmu()
pol = Policy(mode=1, mms=2, khits=1)
it = []
CH = 50_000
for a in range(0, n, CH):
    m = min(CH, n - a)
    codes, quals, offs, seeds, _ = bench.make_reads(genome, m, seed=(777, a // CH))

    class B:
        def __len__(self):
            return m
    b = B()
    b.seq_codes, b.qual_cat, b.offs, b.seeds = codes, quals, offs, seeds
    emu.align(base, b, pol)
    it.append(np.array(emu.iters_per_read, np.int64))
    print(a + m, "reads", file=sys.stderr)
it = np.concatenate(it)
tot = int(it.sum())
res = {"index": name, "reads": int(n), "mean": float(it.mean()),
       "quantiles": {str(q): int(np.quantile(it, q)) for q in (0.5, 0.9, 0.99, 0.999, 0.9999)}, "max": int(it.max()),
       "reads_over": {str(t): int((it > t).sum()) for t in (2000, 4000, 8000, 16000, 32000, 65536, 131072, 262144, 524288)},
       "share_of_transitions_over": {str(t): float(it[it > t].sum() / tot) for t in (8000, 32000, 131072, 262144)},
       "main_budget_8000": {"transitions_in_main": int(np.minimum(it, 8000).sum()), "transitions_in_tail_rerun": int(it[it > 8000].sum())}}
# the restart tail, modelled: over-budget reads in arrival order, 32 (one warp) or 128 (one block) per group, every read on its own lane —
# a group stays resident until its longest read ends; "dense" is the same lane-time if every warp were always full
t = it[it > 8000]
for g in (32, 128):
    k = (len(t) // g) * g
    if k:
        grp = t[:k].reshape(-1, g)
        res[f"tail_model_groups_of_{g}"] = {"groups": int(grp.shape[0]), "sum_of_group_maxima": int(grp.max(axis=1).sum()) * (g // 32),
                                            "dense_equivalent": int(t[:k].sum() // 32), "ratio": float(grp.max(axis=1).sum() * (g // 32) / max(1, t[:k].sum() // 32))}
print(json.dumps(res, indent=1))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(json.dumps(res, indent=1) + "\n")
