#!/usr/bin/env python
"""Development micro-measurements on the GPU box (not the bench): lone-lane latency of the heaviest reads,
kernel throughput on the tail-free e_coli index, and step time on the bench index at several batch sizes."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import bowtie_b200  # noqa: E402


def dev_batch(h):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (h[0], h[1], h[2].view(np.int64), h[3].view(np.int32)))


def time_align(ix, dev, n, pol, reps=3, mm_cap=7):
    rw = 5 + mm_cap
    f = torch.zeros(n, dtype=torch.int32, device="cuda"); g = torch.zeros(n, dtype=torch.int32, device="cuda")
    h = torch.zeros(n * rw, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    best = 1e30
    for r in range(reps + 1):
        ix.stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.align_device(dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), n, 100, pol, f.data_ptr(), g.data_ptr(), h.data_ptr(), 1, mm_cap, st)
        e1.record(); torch.cuda.synchronize()
        if r > 0:
            best = min(best, e0.elapsed_time(e1))
    s = ix.stats(reset=True)
    return best, s, int((f > 0).sum().item()), int((g != 0).sum().item())


def main():
    if not os.environ.get('BOWTIE_B200_LIB'):
        bowtie_b200.build_library()
    pol = bowtie_b200.Policy(mode=1, mms=2)
    which = sys.argv[1:] or ["lone", "ecoli", "bench"]
    base, name = bench.pick_index()
    if "lone" in which or "bench" in which:
        ix = bowtie_b200.Index(str(base))
        genome = bench.load_genome(base)
    if "lone" in which:
        hv = json.loads((ROOT / "tools" / "heavy_reads.json").read_text())
        codes, quals, offs, seeds, nm = bench.make_reads(genome, hv["n"], hv["seed"])
        for k in (0,):
            rid = hv["read_ids"][k]
            one = (codes[rid * 100:(rid + 1) * 100].copy(), quals[rid * 100:(rid + 1) * 100].copy(), np.array([0, 100], np.uint64), seeds[rid:rid + 1].copy())
            ms, s, al, fl = time_align(ix, dev_batch(one), 1, pol)
            print(json.dumps({"test": "lone", "read": rid, "emu_iters": hv["iters"][k], "gpu_iters": s.iters, "ms": ms, "us_per_iter": 1e3 * ms / max(1, s.iters), "flags": fl}))
        # the 20 heaviest together, and a batch of 20k ordinary reads
        ids = hv["read_ids"]
        cc = np.concatenate([codes[r * 100:(r + 1) * 100] for r in ids]); qq = np.concatenate([quals[r * 100:(r + 1) * 100] for r in ids])
        many = (cc, qq, (np.arange(len(ids) + 1) * 100).astype(np.uint64), seeds[ids].copy())
        ms, s, al, fl = time_align(ix, dev_batch(many), len(ids), pol)
        print(json.dumps({"test": "heavy20", "ms": ms, "iters": s.iters, "flags": fl}))
    if "bench" in which:
        for n in (1_000_000,):
            h = bench.make_reads(genome, n, 777)
            ms, s, al, fl = time_align(ix, dev_batch(h), n, pol, reps=2)
            print(json.dumps({"test": "bench_index", "n": n, "ms": ms, "reads_per_s": n / ms * 1e3, "iters_per_read": s.iters / n, "lane_iters_per_s": s.iters / ms * 1e3,
                              "aligned": al / n, "flags": fl, "side_fetches_per_read": s.side_fetches / n}))
    if "ecoli" in which:
        eb = ROOT / "oracle" / "_ref" / "fixtures" / "e_coli"
        ix2 = bowtie_b200.Index(str(eb))
        g2 = bench.load_genome(eb)
        for n in (2_000_000,):
            h = bench.make_reads(g2, n, 778)
            for p, nm in ((pol, "-n 2"), (bowtie_b200.Policy(mode=0, mms=0), "-v 0")):
                ms, s, al, fl = time_align(ix2, dev_batch(h), n, p, reps=2)
                print(json.dumps({"test": "ecoli", "policy": nm, "n": n, "ms": ms, "reads_per_s": n / ms * 1e3, "iters_per_read": s.iters / n,
                                  "lane_iters_per_s": s.iters / ms * 1e3, "aligned": al / n, "flags": fl, "GBps_alg": s.algorithmic_bytes / ms / 1e6}))


if __name__ == "__main__":
    main()
