"""Logic check of the DEVICE state machine (bowtie_b200/csrc/bt_core.cuh) without a GPU: the header is
compiled for the host by tests/host_emu/emu.cpp (test-only) and compared with the oracle, including the
operation counters that feed the roofline accounting.  The real kernels are checked in test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

from helpers import HostEmu, Policy, results_equal

POLICIES = [
    Policy(mode=0, mms=0), Policy(mode=0, mms=1), Policy(mode=0, mms=2), Policy(mode=0, mms=2, all_hits=True),
    Policy(mode=1, mms=0), Policy(mode=1, mms=1), Policy(mode=1, mms=2), Policy(mode=1, mms=3),
    Policy(mode=1, mms=2, all_hits=True), Policy(mode=1, mms=2, mhits=1), Policy(mode=1, mms=2, khits=3),
    Policy(mode=1, mms=2, nofw=True), Policy(mode=1, mms=3, norc=True, khits=2),
    Policy(mode=1, mms=2, maq_round=False, qual_thresh=100), Policy(mode=1, mms=2, seed_len=20, max_bts=10),
]


@pytest.fixture(scope="module")
def emu():
    return HostEmu()


def test_relayout_lf_matches_side_arithmetic(emu, oracle, ecoli_base):
    """LF on the 32-byte rank blocks == Ebwt::mapLFEx / rowL on the native sides, fw and mirror index."""
    rng = np.random.default_rng(5)
    for mirror in (False, True):
        e = emu.index(ecoli_base, mirror)
        o = oracle.index(ecoli_base, mirror)
        length = 4938920
        rows = np.concatenate([rng.integers(0, length + 1, 20000), np.arange(0, 300), np.arange(length - 300, length + 1),
                               np.arange(780711 - 70, 780711 + 70)])
        for r in rows.tolist():
            a = (C.c_uint32 * 4)(); b = (C.c_uint32 * 4)()
            oracle.L.bto_map_lf_ex(o, r, a)
            emu.L.emu_lf_ex(e, r, b)
            assert list(a) == list(b), (mirror, r)
            if r < length + 1:
                assert oracle.L.bto_row_l(o, r) == emu.L.emu_row_l(e, r)
                c = oracle.L.bto_row_l(o, r)
                assert oracle.L.bto_map_lf(o, r, c) == emu.L.emu_lf(e, r, c)


@pytest.mark.parametrize("pol", POLICIES, ids=lambda p: " ".join(p.ref_args()))
def test_state_machine_matches_oracle_ecoli(pol, emu, oracle, ecoli_base, ecoli_reads):
    a = oracle.align(ecoli_base, ecoli_reads, pol)
    b, flags = emu.align(ecoli_base, ecoli_reads, pol)
    assert not flags.any()
    ok, why = results_equal(a, b)
    assert ok, why
    for k in ("lfex", "lf", "chase", "ftab", "offs", "backtracks"):
        assert a.stats[k] == b.stats[k], k


@pytest.mark.parametrize("pol", POLICIES[::2], ids=lambda p: " ".join(p.ref_args()))
def test_state_machine_matches_oracle_synthetic(pol, emu, oracle, synth_index):
    """Multi-reference index with N gaps (nFrag > nPat), ragged read lengths, Ns, low qualities (deep recursion)."""
    from synth import synth_reads
    base, genome = synth_index
    batch = synth_reads(genome, 600, (18, 120), seed=3, sub_rate=0.03, n_rate=0.005, qual_profile="low")
    a = oracle.align(base, batch, pol)
    b, flags = emu.align(base, batch, pol, mm_cap=64, FCAP=128, PCAP=4096, R=120 * 121)
    assert not flags.any()
    ok, why = results_equal(a, b)
    assert ok, why


def test_scratch_overflow_is_flagged_not_silent(emu, oracle, synth_index):
    """With a deliberately tiny workspace the lane must flag the read (the library then retries it)."""
    from synth import synth_reads
    base, genome = synth_index
    batch = synth_reads(genome, 300, 80, seed=9, sub_rate=0.04, qual_profile="low")
    pol = Policy(mode=1, mms=3)
    a = oracle.align(base, batch, pol)
    b, flags = emu.align(base, batch, pol, mm_cap=2, FCAP=2, PCAP=2, R=90)
    assert flags.any()
    clean = flags == 0
    # reads that were not flagged are exact
    assert np.array_equal(a.nhits_per_read[clean], b.nhits_per_read[clean])


@pytest.mark.parametrize("pol", POLICIES, ids=lambda p: " ".join(p.ref_args()))
def test_time_sliced_search_equals_oracle(pol, emu, oracle, ecoli_base, ecoli_reads):
    """Checkpoint slots (bt_ctxq.cuh): every read is suspended after a few transitions — packed lane state, read copy and live scratch
    into a slot, the lane poisoned — and resumed from the slot, over and over with a growing budget.  Output and operation counters
    equal the oracle's, i.e. a suspended read loses nothing."""
    a = oracle.align(ecoli_base, ecoli_reads, pol)
    for budget0, growth in ((7, 2), (150, 3)):
        b, flags, nsusp = emu.align_sliced(ecoli_base, ecoli_reads, pol, budget0, growth)
        assert budget0 > 7 or nsusp > 0          # (the budget is tested between rare transitions: exact-match searches have few)
        assert not flags.any()
        ok, why = results_equal(a, b)
        assert ok, (budget0, growth, why)
        for k in ("lfex", "lf", "chase", "ftab", "offs", "backtracks"):
            assert a.stats[k] == b.stats[k], k


@pytest.mark.parametrize("pol", POLICIES[::2], ids=lambda p: " ".join(p.ref_args()))
def test_time_sliced_search_synthetic(pol, emu, oracle, synth_index):
    """Ragged reads, Ns, deep recursion and long seedling lists: the main pass's small scratch (8 frames, 16 seedlings here) fills up, the
    read moves to a slot with the later passes' capacities and finishes there."""
    from synth import synth_reads
    base, genome = synth_index
    batch = synth_reads(genome, 600, (18, 120), seed=3, sub_rate=0.03, n_rate=0.005, qual_profile="low")
    a = oracle.align(base, batch, pol)
    b, flags, nsusp = emu.align_sliced(base, batch, pol, 40, 2, mm_cap=64, R=120 * 121, FCAP=128, PCAP=16, slot_FCAP=128, slot_PCAP=4096)
    assert nsusp > 100
    assert not flags.any()
    ok, why = results_equal(a, b)
    assert ok, why


@pytest.mark.parametrize("pol", [Policy(mode=1, mms=2), Policy(mode=0, mms=2), Policy(mode=1, mms=3, khits=2, maq_round=False, qual_thresh=120)], ids=lambda p: " ".join(p.ref_args()))
def test_long_reads_live_mask_rows(pol, emu, oracle, synth_index):
    """Reads of 200-600 bases: every frame's live-position mask (bt_live_mask, one row per 256 positions below the frame's rowbase) spans
    several rows, and the backtrack-target scan / next-lowest-quality re-scan walk its words from the top — output and operation
    counters equal the oracle's position-by-position loops."""
    from synth import synth_reads
    base, genome = synth_index
    batch = synth_reads(genome, 120, (200, 600), seed=21, sub_rate=0.02, n_rate=0.002, qual_profile="low")
    a = oracle.align(base, batch, pol)
    b, flags = emu.align(base, batch, pol, mm_cap=96, FCAP=64, PCAP=4096, R=60000)
    assert not flags.any()
    ok, why = results_equal(a, b)
    assert ok, why
    for k in ("lfex", "lf", "chase", "ftab", "offs", "backtracks"):
        assert a.stats[k] == b.stats[k], k
