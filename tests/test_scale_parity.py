"""Parity at BASELINE scale: the workloads bench.py times — 100-bp reads / 2x100-bp pairs sampled from the
synthetic repeat-bearing genomes, `-n 2 -k 1`, `-n 2 --best`, paired `-n 3` — compared record by record (strand, reference, offset,
other-matches count, mismatch list) with the unmodified reference binary on the same reads and index:

* a 256-Mbp / 24-sequence genome of the same family (the size round 1 measured on);
* the hg19-sized (3.0-Gbp) index bench.py measures on (joined offsets beyond 2^31, nFrag > nPat, reads that straddle fragment
  boundaries; ebwt.h:2569-2629).
Both indexes are built once per box by bt_index_build_text into bench.py's cache directory (seconds on a B200; the files are
byte-identical to bowtie-build's, tests/test_index_build.py).

The comparison code is bench.py's own `parity_sample` machinery, so the bench line's parity claim is the one tested here.
"""
import os
import tempfile
from pathlib import Path

import numpy as np
import pytest

from helpers import ROOT, ensure_oracle_built, have_reference

pytestmark = pytest.mark.gpu



def _check(base: Path, policy: str, n: int, seed: int):
    import bench
    import bowtie_b200
    ensure_oracle_built()
    if not have_reference():
        pytest.skip("reference binary not available")
    bowtie_b200.build_library()
    pd = bench.POLICIES[policy]
    paired = pd["paired"]
    R = 2 if paired else 1
    genome = bench.load_genome(base)
    h = (bench.make_pairs if paired else bench.make_reads)(genome, n, seed=seed)
    del genome
    ix = bowtie_b200.Index(str(base), need_mirror=True, device=0)
    pol = bench.lib_policy(policy)
    found, flags, hits = ix.align(h[0], h[1], h[2], h[3], pol, slots=R, mm_cap=7)
    assert not (flags[: n] != 0).any(), "overflow flags on the bench workload"
    got = bench.gpu_records(found, hits, n, paired)
    with tempfile.TemporaryDirectory() as tdn:
        td = Path(tdn)
        fq = bench.write_sample(td, h, n, paired)
        err: list = []
        bench.run_reference(base, fq, min(32, bench.host_cores()), pd["flags"], td / "ref.out", stderr_to=err)
        ref = bench.parse_reference_output(td / "ref.out", [x.decode() if isinstance(x, bytes) else x for x in ix.refnames])
    ix.close()
    # ChunkPool exhaustion (pool.h:146-165: "Exhausted best-first chunk memory for read ... skipping read") is the one documented
    # deviation of the best-first / paired path: the reference drops the rest of that read's search when its 64-MB
    # pool is used up, this implementation finishes it.  Pinned here at read granularity: every read the reference did NOT give up on
    # must be identical, and the reads it gave up on are few (the paired `-n 3` workload reaches the limit for ~1 pair in 10^4).
    import re
    skipped = {int(m) for m in re.findall(r"Exhausted best-first chunk memory for read \S+ \(patid (\d+)\)", err[0])}
    bad = [k for k in set(ref) | set(got) if ref.get(k) != got.get(k)]
    unexplained = [k for k in bad if k[0] not in skipped]
    assert not unexplained, (len(bad), unexplained[:3], [(ref.get(k), got.get(k)) for k in unexplained[:3]])
    assert len(skipped) <= max(2, n // 2000), f"the reference exhausted its chunk pool for {len(skipped)} of {n} units"
    if policy != "paired":
        assert not skipped, "chunk-pool exhaustion on an unpaired BASELINE configuration"
    assert len(ref) > 0.5 * n * R
    print(f"{policy}: {len(ref)} records identical; reference gave up on {len(skipped)} units (chunk pool), {len(bad)} records differ there")
    return ref


def _index(mbp):
    import bench
    try:
        return bench.ensure_index(mbp, 0)[0]
    except Exception as ex:          # no room / no time on this box: the bench reports the same failure loudly
        pytest.skip(f"{mbp}-Mbp index unavailable: {ex}")


@pytest.fixture(scope="module")
def idx256():
    return _index(256)


@pytest.mark.parametrize("policy,n", [("n2k1", 200_000), ("best", 100_000), ("paired", 100_000)])
def test_bench_workload_matches_reference_256mbp(idx256, policy, n):
    _check(idx256, policy, n, seed=4242)


@pytest.fixture(scope="module")
def hg19_index():
    return _index(int(os.environ.get("BT_TEST_HG19_MBP", 3000)))


@pytest.mark.parametrize("policy,n", [("n2k1", 100_000), ("best", 100_000), ("paired", 100_000)])
def test_bench_workload_matches_reference_hg19_sized(hg19_index, policy, n):
    ref = _check(hg19_index, policy, n, seed=777)
    # the hits must cover the far end of the joined text: sequences whose joined offset lies beyond 2^31 (chr13.. at 3 Gbp)
    import bench
    if bench.index_len(hg19_index) > (1 << 31):
        tidx = np.array([v[1] for v in ref.values()])
        assert (tidx >= 20).any() and (tidx <= 2).any()


def test_tail_modes_agree(idx256):
    """The two ways of finishing the reads over the main pass's budget — re-run from scratch by one pass (default) and the round-robin
    tail over checkpoint slots (BT_TAIL=rr, bt_tail.cu: suspension, ring, resumption on another SM) — must produce the same records.
    A small main budget sends a few per cent of the reads through the tail (BT_MAIN_BUDGET=600; the default 8000 sends 0.7 %)."""
    import subprocess
    import sys
    outs = {}
    with tempfile.TemporaryDirectory() as td:
        for mode, env_extra in (("restart", {"BT_TAIL": "restart"}), ("rr", {"BT_TAIL": "rr", "BT_TAIL_QUANTUM": "512"}), ("rr_dense", {"BT_TAIL": "rr", "BT_TAIL_WTARGET": "16", "BT_TAIL_MINCAP": "32"})):
            env = {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"}
            env.update(env_extra); env["BT_MAIN_BUDGET"] = "600"
            out = str(Path(td) / f"{mode}.npz")
            p = subprocess.run([sys.executable, str(ROOT / "tests" / "tail_mode_child.py"), str(idx256), "150000", "99", out], env=env, capture_output=True, text=True, timeout=900)
            assert p.returncode == 0, p.stderr[-2000:]
            outs[mode] = dict(np.load(out))
    a = outs["restart"]
    assert (a["found"] > 0).mean() > 0.5 and not a["flags"].any()
    for mode in ("rr", "rr_dense"):
        b = outs[mode]
        assert not b["flags"].any()
        assert np.array_equal(a["found"], b["found"]), mode
        rw = a["hits"].size // a["found"].size
        ha, hb = a["hits"].reshape(-1, rw), b["hits"].reshape(-1, rw)
        sel = a["found"] > 0
        assert np.array_equal(ha[sel], hb[sel]), mode
        assert int(b["side"][0]) <= int(a["side"][0])       # nothing is re-run: the round-robin tail never fetches more sides than the restart tail counts
