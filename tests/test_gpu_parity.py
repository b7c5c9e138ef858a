"""Parity tests proper: the CUDA path, called through the C ABI (libbowtie_b200.so), against the oracle
on the same inputs.  Bit-exact: every hit field, mismatch list, per-read counts and the -m/-k semantics."""
import numpy as np
import pytest

from helpers import FIXTURES, ROOT as ROOT_DIR, Policy, decode_device_result, results_equal

pytestmark = pytest.mark.gpu

POLICIES = [
    Policy(mode=0, mms=0), Policy(mode=0, mms=1), Policy(mode=0, mms=2), Policy(mode=0, mms=2, all_hits=True),
    Policy(mode=1, mms=0), Policy(mode=1, mms=1), Policy(mode=1, mms=2), Policy(mode=1, mms=3),
    Policy(mode=1, mms=2, all_hits=True), Policy(mode=1, mms=2, mhits=1), Policy(mode=1, mms=2, khits=3),
    Policy(mode=1, mms=2, nofw=True), Policy(mode=1, mms=3, norc=True, khits=2),
    Policy(mode=1, mms=2, maq_round=False, qual_thresh=100), Policy(mode=1, mms=2, seed_len=20, max_bts=10),
]


def to_dev(p: Policy):
    import bowtie_b200
    return bowtie_b200.Policy(**p.__dict__)


def gpu_align(ix, batch, pol, slots=None, mm_cap=8):
    import bowtie_b200
    if slots is None:
        slots = 64 if pol.all_hits else pol.khits
    found, flags, hits = ix.align(batch.seq_codes, batch.qual_cat, batch.offs, batch.seeds, to_dev(pol), slots=slots, mm_cap=mm_cap)
    # caller-visible capacity overflows: retry those reads with exact capacities (ABI contract)
    need = np.nonzero(flags & (bowtie_b200.OVF_HITS | bowtie_b200.OVF_MM))[0]
    assert not (flags & 7).any(), "scratch overflows must be resolved inside the library"
    if len(need):
        lim = 0xFFFFFFFF if pol.all_hits else pol.khits
        slots2 = int(min(int(found.max()), lim))
        maxlen = int((batch.offs[1:] - batch.offs[:-1]).max())
        f2 = found.copy(); g2 = flags.copy(); h2 = np.zeros((len(found), slots2, 5 + maxlen), np.uint32)
        ix.align(batch.seq_codes, batch.qual_cat, batch.offs, batch.seeds, to_dev(pol), slots=slots2, mm_cap=maxlen,
                 sel=need.astype(np.uint32), out=(f2, g2, h2))
        assert not g2[need].any()
        big = np.zeros((len(found), slots2, 5 + maxlen), np.uint32)
        s0 = min(slots, slots2)
        big[:, :s0, :5 + mm_cap] = hits[:, :s0, :]
        big[need] = h2[need]
        found[need] = f2[need]
        return decode_device_result(found, big.reshape(-1), slots2, maxlen, pol)
    return decode_device_result(found, hits.reshape(-1), slots, mm_cap, pol)


@pytest.fixture(scope="module")
def ecoli_ix(ecoli_base):
    import bowtie_b200
    bowtie_b200.build_library()
    return bowtie_b200.Index(str(ecoli_base), need_mirror=True)


@pytest.fixture(scope="module")
def synth_ix(synth_index):
    import bowtie_b200
    return bowtie_b200.Index(str(synth_index[0]), need_mirror=True)


def test_device_lf_matches_oracle(ecoli_ix, oracle, ecoli_base):
    import ctypes as C
    rng = np.random.default_rng(1)
    length = ecoli_ix.len
    rows = np.concatenate([rng.integers(0, length + 1, 200000), np.arange(0, 300), np.arange(length - 300, length + 1),
                           np.arange(780711 - 70, 780711 + 70)]).astype(np.uint32)
    for mirror in (False, True):
        got = ecoli_ix.debug_lf(rows, mirror=mirror)
        o = oracle.index(ecoli_base, mirror)
        sub = np.concatenate([np.arange(0, 3000), np.arange(len(rows) - 800, len(rows))])
        for i in sub.tolist():
            a = (C.c_uint32 * 4)()
            oracle.L.bto_map_lf_ex(o, int(rows[i]), a)
            assert list(a) == got[i, :4].tolist(), (mirror, int(rows[i]))
            assert oracle.L.bto_row_l(o, int(rows[i])) == int(got[i, 4])
        assert not (got[:, :4] == 0xDEADBEEF).any()


@pytest.mark.parametrize("pol", POLICIES, ids=lambda p: " ".join(p.ref_args()))
def test_gpu_matches_oracle_ecoli(pol, ecoli_ix, oracle, ecoli_base, ecoli_reads):
    a = oracle.align(ecoli_base, ecoli_reads, pol)
    b = gpu_align(ecoli_ix, ecoli_reads, pol)
    ok, why = results_equal(a, b)
    assert ok, why


def test_gpu_op_counters_match_oracle(ecoli_ix, oracle, ecoli_base, ecoli_reads):
    """The side-fetch counters that feed roofline.achieved are the oracle's (SURVEY.md §8d units)."""
    pol = Policy(mode=1, mms=2)
    ecoli_ix.stats(reset=True)
    gpu_align(ecoli_ix, ecoli_reads, pol)
    s = ecoli_ix.stats(reset=True)
    a = oracle.align(ecoli_base, ecoli_reads, pol)
    assert (s.lfex, s.lf, s.chase, s.ftab, s.offs) == tuple(a.stats[k] for k in ("lfex", "lf", "chase", "ftab", "offs"))


@pytest.mark.parametrize("pol", POLICIES, ids=lambda p: " ".join(p.ref_args()))
def test_gpu_matches_oracle_synthetic_ragged(pol, synth_ix, oracle, synth_index):
    from synth import synth_reads
    base, genome = synth_index
    batch = synth_reads(genome, 3000, (18, 120), seed=3, sub_rate=0.03, n_rate=0.005, qual_profile="low")
    a = oracle.align(base, batch, pol)
    b = gpu_align(synth_ix, batch, pol)
    ok, why = results_equal(a, b)
    assert ok, why


def test_gpu_matches_golden_md5(ecoli_ix, ecoli_reads, ecoli_base):
    """End-to-end against the committed golden vectors of the reference binary (tests/golden/)."""
    import json
    from helpers import GOLDEN, load_refnames, md5, render_default
    cases = json.loads((GOLDEN / "ecoli_golden.json").read_text())
    names = load_refnames(ecoli_base)
    for case in cases:
        pol = Policy(**case["policy"])
        res = gpu_align(ecoli_ix, ecoli_reads, pol)
        assert md5(render_default(ecoli_reads, res, names)) == case["md5"], case["flags"]


def test_gpu_empty_and_tiny_reads(ecoli_ix, oracle, ecoli_base):
    from helpers import finalize_batch
    batch = finalize_batch([b"e0", b"e1", b"e2", b"e3", b"e4"], [b"", b"A", b"ACG", b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", b"ACGTACGTAC"],
                           [b"", b"I", b"III", b"I" * 30, b"I" * 10])
    for pol in (Policy(mode=0, mms=0), Policy(mode=1, mms=2)):
        a = oracle.align(ecoli_base, batch, pol)
        b = gpu_align(ecoli_ix, batch, pol)
        ok, why = results_equal(a, b)
        assert ok, why


def test_gpu_large_batch_properties(ecoli_ix, oracle, ecoli_base):
    """Size-independent properties at a batch far beyond what the oracle is timed on:
       idempotence (same batch twice), partition invariance (split batches == whole batch),
       and exact agreement with the oracle on a random sample."""
    from synth import synth_reads
    genome = [(">e", open(FIXTURES / "NC_008253.fna", "rb").read().split(b"\n", 1)[1].replace(b"\n", b""))]
    batch = synth_reads(genome, 200_000, 100, seed=12345, sub_rate=0.01, qual_profile="mixed")
    pol = Policy(mode=1, mms=2)
    r1 = gpu_align(ecoli_ix, batch, pol)
    r2 = gpu_align(ecoli_ix, batch, pol)
    ok, why = results_equal(r1, r2)
    assert ok, why
    from helpers import ReadBatch
    half = len(batch) // 2
    def sub(lo, hi):
        o = batch.offs[lo:hi + 1] - batch.offs[lo]
        return ReadBatch(batch.names[lo:hi], batch.seqs[lo:hi], batch.quals[lo:hi], batch.seq_codes[int(batch.offs[lo]):int(batch.offs[hi])],
                         batch.qual_cat[int(batch.offs[lo]):int(batch.offs[hi])], o.astype(np.uint64), batch.seeds[lo:hi])
    ra = gpu_align(ecoli_ix, sub(0, half), pol)
    rb = gpu_align(ecoli_ix, sub(half, len(batch)), pol)
    assert np.array_equal(np.concatenate([ra.nhits_per_read, rb.nhits_per_read]), r1.nhits_per_read)
    assert np.array_equal(np.concatenate([ra.hits["toff"], rb.hits["toff"]]), r1.hits["toff"])
    samp = sub(1000, 6000)
    a = oracle.align(ecoli_base, samp, pol)
    b = gpu_align(ecoli_ix, samp, pol)
    ok, why = results_equal(a, b)
    assert ok, why
    assert r1.counters[0] > 0.9 * len(batch)


def test_gpu_concurrent_contexts_match_single(ecoli_ix, oracle, ecoli_base):
    """Four batches in flight on four contexts / CUDA streams (the pipelined form bench.py uses) give exactly the
    results of the synchronous entry point, and reads that exceed the main-pass budget come back through the
    heavy pass with the same answers."""
    import torch
    import bowtie_b200
    from synth import synth_reads
    genome = [(">e", open(FIXTURES / "NC_008253.fna", "rb").read().split(b"\n", 1)[1].replace(b"\n", b""))]
    pol = Policy(mode=1, mms=2, khits=2)
    dpol = to_dev(pol)
    batches = [synth_reads(genome, 20000, 100, seed=100 + i, sub_rate=0.02, qual_profile="mixed") for i in range(4)]
    want = [gpu_align(ecoli_ix, b, pol, slots=2, mm_cap=8) for b in batches]
    ctxs = [bowtie_b200.Context(ecoli_ix) for _ in range(4)]
    streams = [torch.cuda.Stream() for _ in range(4)]
    dev, outs = [], []
    for b in batches:
        dev.append([torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (b.seq_codes, b.qual_cat, b.offs.view(np.int64), b.seeds.view(np.int32))])
        n = len(b)
        outs.append((torch.zeros(n, dtype=torch.int32, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda"),
                     torch.zeros(n * 2 * 13, dtype=torch.int32, device="cuda")))
    torch.cuda.synchronize()
    for rep in range(2):
        for i in range(4):
            d, o = dev[i], outs[i]
            ctxs[i].align_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), len(batches[i]), 100, dpol,
                                 o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), 2, 8, streams[i].cuda_stream)
    for i in range(4):
        ctxs[i].join(streams[i].cuda_stream)
    torch.cuda.synchronize()
    for i in range(4):
        found = outs[i][0].cpu().numpy().view(np.uint32)
        flags = outs[i][1].cpu().numpy().view(np.uint32)
        hits = outs[i][2].cpu().numpy().view(np.uint32)
        assert not flags.any()
        got = decode_device_result(found, hits, 2, 8, pol)
        ok, why = results_equal(want[i], got)
        assert ok, (i, why)
    for c in ctxs:
        c.close()


def test_gpu_heavy_pass_is_exact(ecoli_base, oracle, ecoli_reads):
    """With a main-pass budget of 40 transitions nearly every read is finished by the heavy pass; results unchanged.
    (Runs in a subprocess because the budget is read once per process.)"""
    import subprocess, sys, json
    code = """
import sys, json, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import bowtie_b200
from helpers import parse_fastq, FIXTURES, Policy, decode_device_result, md5, render_default, load_refnames
ix = bowtie_b200.Index(str(FIXTURES / 'e_coli'))
b = parse_fastq(FIXTURES / 'e_coli_1000.fq')
pol = Policy(mode=1, mms=2)
f, g, h = ix.align(b.seq_codes, b.qual_cat, b.offs, b.seeds, bowtie_b200.Policy(**pol.__dict__), slots=1, mm_cap=8)
assert not g.any()
res = decode_device_result(f, h.reshape(-1), 1, 8, pol)
print(md5(render_default(b, res, load_refnames(FIXTURES / 'e_coli'))))
""" % (str(ROOT_DIR), str(ROOT_DIR / "tests"))
    import os
    env = dict(os.environ, BT_MAIN_BUDGET="40")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    assert p.stdout.strip() == "7238b0f529dfcdf602f82ae1a754a1bd"      # golden md5 of the reference binary for -n 2
