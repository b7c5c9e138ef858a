#!/usr/bin/env python
"""Mutation fuzzer for the driver's read sources: takes well-formed FASTQ / FASTA / raw / --12 input, damages a few bytes (newlines
added or removed, CRs, stray '>' '@' '+', spaces, truncation) and compares output, exit status and messages with the reference
binary.  Inputs on which the reference is not deterministic (it parses leftovers of its buffers there) or crashes are skipped.
Usage: python tools/fuzz_input.py [--iters 300] [--seed 1]"""
import argparse
import os
import random
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
FIX = REF / "fixtures"
CLI = Path(os.environ.get("BT_FUZZ_CLI", ROOT / "bowtie_b200" / "bowtie-b200-align"))
SHIM = Path(os.environ.get("BT_FUZZ_SHIM", ROOT / "tests" / "host_emu" / "shim"))


def run(exe, flags, src, out, env=None, extra=()):
    try:
        p = subprocess.run([str(exe), *flags, *extra, "-x", str(FIX / "e_coli"), *src, str(out)], capture_output=True, text=True, env=env, errors="replace", timeout=60)
    except subprocess.TimeoutExpired:
        return -999, b"", ["(timed out)"]
    body = Path(out).read_bytes() if p.returncode == 0 and Path(out).exists() else b""
    return p.returncode, body, [l for l in p.stderr.splitlines() if not l.startswith("Command:")]


def mutate(rng, data: bytes) -> bytes:
    b = bytearray(data)
    for _ in range(rng.randint(1, 3)):
        if not b:
            break
        pos = rng.randrange(len(b))
        what = rng.randrange(9)
        if what == 0:
            b.insert(pos, 10)
        elif what == 1:
            b[pos:pos + 1] = b"\r\n"
        elif what == 2:
            nl = b.find(b"\n", pos)
            if nl >= 0:
                del b[nl]
        elif what == 3:
            b.insert(pos, rng.choice(b">@+ \t.-N"))
        elif what == 4:
            del b[pos]
        elif what == 5:
            del b[pos:]
        elif what == 6:
            b[pos] = rng.choice(b"acgtnACGTN.-")
        elif what == 7:
            b += b"\n" * rng.randint(1, 2)
        else:
            b.insert(pos, rng.choice(b"\x01\x7f;1"))
    return bytes(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM))
    lines = (FIX / "e_coli_1000.fq").read_text().splitlines()
    recs = [lines[i:i + 4] for i in range(0, len(lines), 4)]
    nfail = nskip = 0
    for it in range(a.iters):
        rng = random.Random(a.seed * 7919 + it)
        n = rng.choice([3, 15, 16, 17, 33])
        rs = [recs[rng.randrange(len(recs))] for _ in range(n)]
        # ("iq", --integer-quals, is left out of the draw: that branch of the reference neither trims nor checks the number of
        # qualities, so a damaged record comes out with a quality string of another length than its sequence)
        kind = rng.choice(["fq", "fq", "fa", "raw", "tab", "il", "pair", "tab5", "fc"])
        if kind == "fq":
            text, opt = "".join("\n".join(r) + "\n" for r in rs), ["-q"]
        elif kind == "fa":
            text, opt = "".join(f">{r[0][1:]}\n{r[1]}\n" for r in rs), ["-f"]
        elif kind == "raw":
            text, opt = "".join(r[1] + "\n" for r in rs), ["-r"]
        elif kind == "fc":                                      # windows of a FASTA input
            text, opt = "".join(f">{r[0][1:]} x\n{r[1]}\n{r[1][::-1]}\n" for r in rs[:6]), ["-F", rng.choice(["20,5", "30,1", "12,7"])]
        elif kind == "iq":
            text, opt = "".join(f"{r[0]}\n{r[1]}\n+\n{' '.join(str(ord(c) - 33) for c in r[3])}\n" for r in rs), ["-q", "--integer-quals"]
        elif kind == "tab":
            text, opt = "".join(f"{r[0][1:]}\t{r[1]}\t{r[3]}\n" for r in rs), ["--12"]
        elif kind == "tab5":
            text, opt = "".join(f"{r[0][1:]}\t{r[1]}\t{r[3]}\t{q[1]}\t{q[3]}\n" for r, q in zip(rs, reversed(rs))), ["--12"]
        else:                                                   # il / pair: FASTQ text, used as --interleaved or as the -1 file
            text, opt = "".join("\n".join(r) + "\n" for r in rs), []
        data = mutate(rng, text.encode())
        flags = rng.choice([["-n", "2"], ["-v", "1"], ["-n", "2", "--best"], ["-n", "2", "-5", "2"], ["-v", "0", "-3", "3"], ["-n", "2", "-s", "3"], ["-v", "1", "-u", "7"],
                            ["-n", "2", "-s", "14", "-u", "5"]])
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            f = td / "in.txt"
            f.write_bytes(data)
            if kind == "il":
                src = ["--interleaved", str(f)]
            elif kind == "pair":
                g = td / "mate2.fq"
                g.write_text("".join("\n".join(r) + "\n" for r in reversed(rs)))
                src = ["-1", str(f), "-2", str(g)] if rng.random() < 0.5 else ["-1", str(g), "-2", str(f)]
            else:
                src = [*opt, str(f)] if not kind.startswith("tab") else ["--12", str(f)]
            r1 = run(REF / "bowtie-align-s", flags, src, td / "r1.out", extra=["-p", "1"])
            r2 = run(REF / "bowtie-align-s", flags, src, td / "r2.out", extra=["-p", "1"])
            if r1 != r2 or r1[0] not in (0, 1):
                nskip += 1
                continue
            o = run(CLI, flags, src, td / "o.out", env=env)
            both_past_end = o[0] == 1 == r1[0] and o[2][:1] != r1[2][:1] and all(m and m[0].startswith("Saw ASCII character") for m in (o[2], r1[2]))
            if both_past_end:                                       # the reference read past the end of a record there: the character it reports is whatever lay behind it
                nskip += 1
                continue
            if o != r1:
                nfail += 1
                keep = Path(f"/tmp/fuzz_input_fail_{a.seed}_{it}")
                subprocess.run(["rm", "-rf", str(keep)]); subprocess.run(["cp", "-r", str(td), str(keep)])
                (keep / "info.txt").write_text(f"{flags} {src}\nref rc {r1[0]} {r1[2][:4]}\nour rc {o[0]} {o[2][:4]}\n")
                print(f"FAIL iter {it}: {kind} {' '.join(flags)} rc ref {r1[0]} ours {o[0]} -> {keep}", flush=True)
        if it % 50 == 49:
            print(f"iter {it + 1}: {nfail} failures, {nskip} skipped", flush=True)
    print(f"done: {a.iters} iterations, {nfail} failures, {nskip} skipped")
    sys.exit(1 if nfail else 0)


if __name__ == "__main__":
    main()
