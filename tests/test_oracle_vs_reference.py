"""Pins the oracle restatement (oracle/bt_oracle.c) to the reference:
   * golden md5s of the reference's own outputs on its shipped fixtures, committed in
     tests/golden/ecoli_golden.json together with the script that generated them;
   * live differential runs against oracle/_ref/bowtie-align-s on synthetic indexes / reads.
"""
import json

import pytest

from helpers import (GOLDEN, Policy, have_reference, load_refnames, md5, render_default, run_reference)

CASES = json.loads((GOLDEN / "ecoli_golden.json").read_text())


def _policy(d):
    return Policy(**d)


@pytest.mark.parametrize("case", CASES, ids=[c["flags"] for c in CASES])
def test_oracle_matches_golden_md5(case, oracle, ecoli_base, ecoli_reads):
    pol = _policy(case["policy"])
    res = oracle.align(ecoli_base, ecoli_reads, pol)
    txt = render_default(ecoli_reads, res, load_refnames(ecoli_base))
    assert md5(txt) == case["md5"]
    assert int(res.counters[0]) + int(res.counters[2]) == case["aligned"]   # stderr counts -m suppressed reads as aligned (hit.h:303-312)
    assert len(txt.splitlines()) == case["lines"]


@pytest.mark.parametrize("pol", [
    Policy(mode=0, mms=0), Policy(mode=0, mms=1, khits=2), Policy(mode=0, mms=2, all_hits=True),
    Policy(mode=1, mms=0), Policy(mode=1, mms=1, mhits=3), Policy(mode=1, mms=2), Policy(mode=1, mms=3, khits=4),
    Policy(mode=1, mms=2, seed_len=20, qual_thresh=100), Policy(mode=1, mms=2, maq_round=False, qual_thresh=90),
    Policy(mode=1, mms=2, nofw=True), Policy(mode=1, mms=3, norc=True), Policy(mode=1, mms=2, max_bts=10),
], ids=lambda p: " ".join(p.ref_args()))
def test_oracle_matches_live_reference_on_synthetic(pol, oracle, synth_index, tmp_path):
    if not have_reference():
        pytest.skip("reference binary not available")
    from synth import synth_reads, write_fastq
    base, genome = synth_index
    batch = synth_reads(genome, 400, (20, 75), seed=11, sub_rate=0.03, n_rate=0.004, qual_profile="low")
    fq = tmp_path / "r.fq"
    write_fastq(fq, batch)
    ref_out, _ = run_reference(pol.ref_args(), base, fq)
    res = oracle.align(base, batch, pol)
    assert render_default(batch, res, load_refnames(base)) == ref_out
