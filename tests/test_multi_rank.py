"""N > 1 path on CPU (gloo, world_size 2): reads shard by contiguous id ranges, every rank searches its shard
independently (here through the host emulation of the device state machine), per-rank results concatenated in
rank order equal the single-rank result (partition invariance: all randomness derives from Read::seed), and the
all-reduced counters equal the whole-batch counters."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, base, out_dir):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    from bowtie_b200.shard import allreduce_counters, counters_from_found, shard_range
    from helpers import HostEmu, Policy, ReadBatch, parse_fastq, FIXTURES
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = parse_fastq(FIXTURES / "e_coli_1000.fq")
    lo, hi = shard_range(len(batch), rank, world)
    o = batch.offs[lo:hi + 1] - batch.offs[lo]
    sub = ReadBatch(batch.names[lo:hi], batch.seqs[lo:hi], batch.quals[lo:hi], batch.seq_codes[int(batch.offs[lo]):int(batch.offs[hi])],
                    batch.qual_cat[int(batch.offs[lo]):int(batch.offs[hi])], o.astype(np.uint64), batch.seeds[lo:hi])
    pol = Policy(mode=1, mms=2, khits=2, mhits=5)
    res, flags = HostEmu().align(base, sub, pol)
    found = res.nhits_per_read.astype(np.int64)      # reported per read; maxed reads are 0 here
    ctr = res.counters.astype(np.int64)
    tot = allreduce_counters(ctr)
    np.save(Path(out_dir) / f"toff_{rank}.npy", res.hits["toff"])
    np.save(Path(out_dir) / f"nhits_{rank}.npy", res.nhits_per_read)
    np.save(Path(out_dir) / f"ctr_{rank}.npy", tot)
    chk = counters_from_found(np.where(res.maxed > 0, 6, found), 2, 5, False)
    assert np.array_equal(chk, ctr), (chk, ctr)
    dist.destroy_process_group()


def test_two_ranks_match_single_rank(ecoli_base, tmp_path):
    import torch.multiprocessing as mp
    from helpers import HostEmu, Policy, parse_fastq, FIXTURES
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(ecoli_base), str(tmp_path)), nprocs=2, join=True)
    batch = parse_fastq(FIXTURES / "e_coli_1000.fq")
    whole, _ = HostEmu().align(ecoli_base, batch, Policy(mode=1, mms=2, khits=2, mhits=5))
    toff = np.concatenate([np.load(tmp_path / f"toff_{r}.npy") for r in range(2)])
    nh = np.concatenate([np.load(tmp_path / f"nhits_{r}.npy") for r in range(2)])
    assert np.array_equal(nh, whole.nhits_per_read)
    assert np.array_equal(toff, whole.hits["toff"])
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"ctr_{r}.npy"), whole.counters.astype(np.int64))


def test_shard_ranges_partition():
    from bowtie_b200.shard import shard_range
    for n in (0, 1, 7, 1000, 12345):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
