"""Generates tests/golden/ecoli_golden.json by running the UNMODIFIED reference binary
(oracle/_ref/bowtie-align-s, built from /root/reference by oracle/Makefile) on the reference's own
shipped fixtures (indexes/e_coli + reads/e_coli_1000.fq).  The md5s equal those listed in
SURVEY.md §8c.  Usage: python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from helpers import FIXTURES, Policy, md5, run_reference  # noqa: E402

CASES = [
    ("-v 0", dict(mode=0, mms=0)), ("-v 1", dict(mode=0, mms=1)), ("-v 2", dict(mode=0, mms=2)),
    ("-n 0", dict(mode=1, mms=0)), ("-n 1", dict(mode=1, mms=1)), ("-n 2", dict(mode=1, mms=2)), ("-n 3", dict(mode=1, mms=3)),
    ("-n 2 -a", dict(mode=1, mms=2, all_hits=True)), ("-n 2 -m 1", dict(mode=1, mms=2, mhits=1)),
    ("-n 2 -k 3", dict(mode=1, mms=2, khits=3)), ("-v 2 -a", dict(mode=0, mms=2, all_hits=True)),
    ("-n 3 -k 2 --norc", dict(mode=1, mms=3, khits=2, norc=True)), ("-n 2 --nofw", dict(mode=1, mms=2, nofw=True)),
    ("-n 2 --nomaqround -e 100", dict(mode=1, mms=2, maq_round=False, qual_thresh=100)),
]
out = []
for flags, pd in CASES:
    pol = Policy(**pd)
    txt, err = run_reference(pol.ref_args(), FIXTURES / "e_coli", FIXTURES / "e_coli_1000.fq")
    aligned = int([l for l in err.splitlines() if "at least one alignment" in l][0].split(":")[1].split()[0])
    out.append({"flags": flags, "policy": pd, "md5": md5(txt), "lines": len(txt.splitlines()), "aligned": aligned})
    print(flags, out[-1]["md5"], out[-1]["lines"], aligned)
(Path(__file__).resolve().parent / "ecoli_golden.json").write_text(json.dumps(out, indent=1) + "\n")
