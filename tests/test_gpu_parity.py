"""Parity tests proper: the CUDA path, called through the C ABI (libbowtie_b200.so), against the oracle
on the same inputs.  Bit-exact: every hit field, mismatch list, per-read counts and the -m/-k semantics."""
import numpy as np
import pytest

from helpers import FIXTURES, Policy, decode_device_result, results_equal

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
