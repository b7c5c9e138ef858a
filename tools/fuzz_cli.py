#!/usr/bin/env python
"""Differential fuzzer in the spirit of the reference's scripts/test/random_bowtie_tests.pl: random small genomes
(repeats, N gaps, several sequences), random reads (ragged lengths, Ns, low qualities) and random option sets for both
search paths and for paired-end alignment; bowtie-b200-align (through tests/host_emu/shim, i.e. the device code compiled for the host — or the real
library with --gpu) must produce the reference binary's hit file and summary byte for byte.

usage: tools/fuzz_cli.py [--iters N] [--seed S] [--gpu]
"""
import argparse
import os
import random
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
CLI = Path(os.environ.get("BT_FUZZ_CLI", ROOT / "bowtie_b200" / "bowtie-b200-align"))      # copies let a long run survive rebuilds
SHIM = Path(os.environ.get("BT_FUZZ_SHIM", ROOT / "tests" / "host_emu" / "shim"))


def rand_genome(rng):
    nseq = rng.randint(1, 4)
    seqs = []
    rep = "".join(rng.choice("ACGT") for _ in range(rng.randint(20, 200)))
    for i in range(nseq):
        L = rng.randint(60, 3000)
        s = [rng.choice("ACGT") for _ in range(L)]
        for _ in range(rng.randint(0, 4)):                       # copies of a repeat, some diverged
            p = rng.randint(0, max(0, L - len(rep)))
            r = list(rep)
            for k in range(len(r)):
                if rng.random() < 0.03:
                    r[k] = rng.choice("ACGT")
            s[p:p + len(r)] = r[:max(0, L - p)]
        if rng.random() < 0.4:                                    # N gap -> more fragments than sequences
            p = rng.randint(0, L - 1)
            for k in range(p, min(L, p + rng.randint(1, 30))):
                s[k] = "N"
        seqs.append(("seq%d some description" % i, "".join(s[:L])))
    return seqs


def rand_reads(rng, genome, n):
    comp = str.maketrans("ACGTN", "TGCAN")
    out = []
    for i in range(n):
        name, g = rng.choice(genome)
        L = rng.randint(4, min(70, len(g)))
        if rng.random() < 0.08:
            r = "".join(rng.choice("ACGT") for _ in range(L))
        else:
            p = rng.randint(0, len(g) - L)
            r = list(g[p:p + L])
            for k in range(L):
                if rng.random() < 0.04:
                    r[k] = rng.choice("ACGT")
                if rng.random() < 0.01:
                    r[k] = "N"
            r = "".join(r)
            if rng.random() < 0.5:
                r = r.translate(comp)[::-1]
        q = "".join(chr(33 + rng.choice([40, 40, 30, 20, 12, 8, 3, 0])) for _ in range(L))
        out.append((f"r{i}", r, q))
    return out


def rand_pairs(rng, genome, n):
    comp = str.maketrans("ACGTN", "TGCAN")
    m1, m2 = [], []
    for i in range(n):
        name, g = rng.choice(genome)
        F = rng.randint(8, min(260, len(g)))
        p = rng.randint(0, len(g) - F)
        frag = g[p:p + F]
        if rng.random() < 0.5:
            frag = frag.translate(comp)[::-1]
        L1, L2 = rng.randint(4, min(60, F)), rng.randint(4, min(60, F))
        a, b = list(frag[:L1]), list(frag[F - L2:].translate(comp)[::-1])
        if rng.random() < 0.1:
            b = [rng.choice("ACGT") for _ in range(L2)]
        for r in (a, b):
            for k in range(len(r)):
                if rng.random() < 0.03:
                    r[k] = rng.choice("ACGT")
                if rng.random() < 0.008:
                    r[k] = "N"
        qa = "".join(chr(33 + rng.choice([40, 40, 30, 20, 12, 8, 3, 0])) for _ in range(L1))
        qb = "".join(chr(33 + rng.choice([40, 40, 30, 20, 12, 8, 3, 0])) for _ in range(L2))
        m1.append((f"p{i}/1", "".join(a), qa)); m2.append((f"p{i}/2", "".join(b), qb))
    return m1, m2


def rand_paired_flags(rng):
    f = []
    if rng.random() < 0.5:
        f += ["-v", str(rng.randint(0, 3))]
    else:
        f += ["-n", str(rng.randint(0, 3)), "-l", str(rng.choice([5, 8, 12, 20, 28])), "-e", str(rng.choice([40, 70, 150, 400]))]
        if rng.random() < 0.3:
            f += ["--nomaqround"]
    rep = rng.choice(["k1", "k", "a", "m", "M"])
    best = rng.random() < 0.4
    if rep == "k":
        f += ["-k", str(rng.randint(2, 5))]
    elif rep == "a":
        f += ["-a"]
    elif rep == "m":
        f += ["-m", str(rng.randint(1, 3))] + rng.choice([[], ["-k", "3"], ["-a"]])
    elif rep == "M":
        f += ["-M", str(rng.randint(1, 3))]
    if best:
        f += ["--best"]                                       # PairedBWAlignerV2
        if rep in ("k", "a", "m") and rng.random() < 0.5:
            f += ["--strata"]
    maxins = 250
    if rng.random() < 0.4:
        maxins = rng.choice([60, 120, 200, 400])
        f += ["-X", str(maxins)]
    if rng.random() < 0.25:
        f += ["-I", str(rng.choice([i for i in (10, 40, 100) if i <= maxins]))]      # -I > -X is undefined behaviour in the reference (it may crash)
    if rng.random() < 0.15:
        f += [rng.choice(["--ff", "--rf", "--fr"])]
    if rng.random() < 0.15:
        f += [rng.choice(["--nofw", "--norc"])]
    if rng.random() < 0.15:
        f += ["--pairtries", str(rng.choice([1, 3, 20]))]
    if rng.random() < 0.15:
        f += ["--maxbts", str(rng.choice([1, 5, 40]))]
    if rng.random() < 0.2:
        f += ["-S"]
    return f


def rand_flags(rng):
    f = []
    if rng.random() < 0.5:
        f += ["-v", str(rng.randint(0, 3))]
    else:
        f += ["-n", str(rng.randint(0, 3)), "-l", str(rng.choice([5, 8, 12, 20, 28, 40])), "-e", str(rng.choice([10, 40, 70, 150, 400]))]
        if rng.random() < 0.3:
            f += ["--nomaqround"]
    best = rng.random() < 0.5
    rep = rng.choice(["k1", "k", "a", "m", "M"])
    if rep == "k":
        f += ["-k", str(rng.randint(2, 6))]
    elif rep == "a":
        f += ["-a"]
    elif rep == "m":
        f += ["-m", str(rng.randint(1, 4))] + rng.choice([[], ["-k", "3"], ["-a"]])
    elif rep == "M":
        f += ["-M", str(rng.randint(1, 4))]
        best = True
    if best:
        f += ["--best"]
        if rep in ("k", "a", "m") and rng.random() < 0.5:
            f += ["--strata"]
    if rng.random() < 0.15:
        f += [rng.choice(["--nofw", "--norc"])]
    if rng.random() < 0.2:
        f += ["--maxbts", str(rng.choice([1, 3, 10, 50]))]
    if rng.random() < 0.1:
        f += ["-y"]
    if rng.random() < 0.2:
        f += ["-S"]
    if rng.random() < 0.2:
        f += ["--seed", str(rng.randint(0, 1000))]
    return f


def rand_io_flags(rng, sam, td):
    """Options of the driver around the search: trimming, read selection, output fields.  Returns (flags, dump_files)."""
    f, dumps = [], []
    if rng.random() < 0.15:
        f += ["-5", str(rng.randint(1, 4))]
    if rng.random() < 0.15:
        f += ["-3", str(rng.randint(1, 4))]
    if rng.random() < 0.1:
        f += ["-s", str(rng.randint(1, 15))]
    if rng.random() < 0.1:
        f += ["-u", str(rng.randint(5, 60))]
    if rng.random() < 0.1:
        f += ["-p", str(rng.randint(2, 4))]                      # ours only formats with threads; the reference side always gets -p 1 last
    if sam:
        if rng.random() < 0.2:
            f += ["--sam-nohead"]
        if rng.random() < 0.2:
            f += ["--sam-nosq"]
        if rng.random() < 0.2:
            f += ["--sam-RG", "ID:x", "--sam-RG", "SM:y"]
        if rng.random() < 0.2:
            f += ["--mapq", str(rng.randint(0, 60))]
        if rng.random() < 0.2:
            f += ["--no-unal"]
    else:
        if rng.random() < 0.15:
            f += ["--refidx"]
        if rng.random() < 0.15:
            f += ["--fullref"]
        if rng.random() < 0.15:
            f += ["-B", str(rng.randint(1, 3))]
        if rng.random() < 0.15:
            f += ["--suppress", ",".join(str(x) for x in sorted(rng.sample(range(1, 9), rng.randint(1, 3))))]
        if rng.random() < 0.1:
            f += ["--cost"]
    if rng.random() < 0.15:
        for opt in rng.sample(["--al", "--un", "--max"], rng.randint(1, 3)):
            f += [opt, f"@DIR@/dump{opt[1:]}.fq"]
            dumps.append(f"dump{opt[1:]}")
    return f, dumps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--keep", default="/tmp/fuzz_fail")
    a = ap.parse_args()
    env = {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"}
    if not a.gpu:
        env["LD_LIBRARY_PATH"] = str(SHIM)
    nfail = 0
    for it in range(a.iters):
        rng = random.Random(a.seed * 100003 + it)
        with tempfile.TemporaryDirectory() as td:
            td = Path(td)
            genome = rand_genome(rng)
            (td / "g.fa").write_text("".join(f">{n}\n{s}\n" for n, s in genome))
            p = subprocess.run([str(REF / "bowtie-build-s"), "-q", "-t", str(rng.choice([1, 2, 4, 6])), "-o", str(rng.choice([1, 3, 5])), str(td / "g.fa"), str(td / "g")],
                               capture_output=True, text=True)
            if p.returncode != 0:
                continue
            reads = rand_reads(rng, [(n, s) for n, s in genome if len(s) >= 4], rng.randint(20, 200))
            (td / "r.fq").write_text("".join(f"@{n}\n{s}\n+\n{q}\n" for n, s, q in reads))
            m1, m2 = rand_pairs(rng, [(n, s) for n, s in genome if len(s) >= 8], rng.randint(20, 150))
            (td / "m1.fq").write_text("".join(f"@{n}\n{s}\n+\n{q}\n" for n, s, q in m1))
            (td / "m2.fq").write_text("".join(f"@{n}\n{s}\n+\n{q}\n" for n, s, q in m2))
            for sub in range(6):
                if sub < 4:
                    flags = rand_flags(rng)
                    kind = rng.choice(["fq", "fq", "fq", "fa", "raw", "tab", "cmd"])           # every read source of the driver
                    if kind == "fq":
                        inputs = [str(td / "r.fq")]
                    elif kind == "fa":
                        (td / "r.fa").write_text("".join(f">{n}\n{s}\n" for n, s, q in reads))
                        inputs = ["-f", str(td / "r.fa")]
                    elif kind == "raw":
                        (td / "r.raw").write_text("".join(f"{s}\n" for n, s, q in reads))
                        inputs = ["-r", str(td / "r.raw")]
                    elif kind == "tab":
                        (td / "r.tab").write_text("".join(f"{n}\t{s}\t{q}\n" for n, s, q in reads))
                        inputs = ["--12", str(td / "r.tab")]
                    else:
                        few = [r for r in reads[:12] if len(r[1]) >= 4 and ":" not in r[2] and "," not in r[2]]
                        if not few:
                            continue
                        inputs = ["-c", ",".join(f"{s}:{q}" for n, s, q in few)]
                else:
                    flags = rand_paired_flags(rng)
                    kind = rng.choice(["fq", "fq", "il", "tab", "fa"])
                    if kind == "fq":
                        inputs = ["-1", str(td / "m1.fq"), "-2", str(td / "m2.fq")]
                    elif kind == "il":
                        (td / "il.fq").write_text("".join(f"@{a[0]}\n{a[1]}\n+\n{a[2]}\n@{b[0]}\n{b[1]}\n+\n{b[2]}\n" for a, b in zip(m1, m2)))
                        inputs = ["--interleaved", str(td / "il.fq")]
                    elif kind == "tab":
                        (td / "p.tab").write_text("".join(f"{a[0][:-2]}\t{a[1]}\t{a[2]}\t{b[1]}\t{b[2]}\n" for a, b in zip(m1, m2)))
                        inputs = ["--12", str(td / "p.tab")]
                    else:
                        (td / "m1.fa").write_text("".join(f">{n}\n{s}\n" for n, s, q in m1))
                        (td / "m2.fa").write_text("".join(f">{n}\n{s}\n" for n, s, q in m2))
                        inputs = ["-f", "-1", str(td / "m1.fa"), "-2", str(td / "m2.fa")]
                io, dumps = rand_io_flags(rng, "-S" in flags, td)
                if dumps and inputs[0] == "-c":                                  # (the reference's -c source never resets its record buffers: its dumps pile up)
                    io = [x for j, x in enumerate(io) if not (x in ("--al", "--un", "--max") or (j > 0 and io[j - 1] in ("--al", "--un", "--max")))]
                flags = flags + io
                (td / "dr").mkdir(exist_ok=True); (td / "do").mkdir(exist_ok=True)
                for d in ("dr", "do"):
                    for x in (td / d).iterdir():
                        x.unlink()
                fr = [x.replace("@DIR@", str(td / "dr")) for x in flags]
                fo = [x.replace("@DIR@", str(td / "do")) for x in flags]
                r = subprocess.run([str(REF / "bowtie-align-s"), *fr, "-p", "1", "-x", str(td / "g"), *inputs, str(td / "ref.out")], capture_output=True, text=True)
                if "Exhausted best-first chunk memory" in r.stderr or r.returncode < 0:     # the reference's own memory limit / a crash of the reference
                    continue
                o = subprocess.run([str(CLI), *fo, "-x", str(td / "g"), *inputs, str(td / "our.out")], capture_output=True, text=True, env=env)
                dump_r = {x.name: x.read_bytes() for x in sorted((td / "dr").iterdir())}
                dump_o = {x.name: x.read_bytes() for x in sorted((td / "do").iterdir())}
                def body(pth):
                    return b"".join(l for l in Path(pth).read_bytes().splitlines(keepends=True) if not l.startswith(b"@PG")) if Path(pth).exists() else b"<none>"
                def summ(t):
                    return [l for l in t.splitlines() if l.startswith("#") or l.startswith("Reported") or l.startswith("No alignments")]
                ok = r.returncode == o.returncode and (r.returncode != 0 or (body(td / "ref.out") == body(td / "our.out") and summ(r.stderr) == summ(o.stderr) and dump_r == dump_o))
                if not ok:
                    nfail += 1
                    keep = Path(a.keep + f"_{a.seed}_{it}_{sub}")
                    subprocess.run(["rm", "-rf", str(keep)]); subprocess.run(["cp", "-r", str(td), str(keep)])
                    (keep / "flags.txt").write_text(" ".join(flags) + "\n" + r.stderr[-2000:] + "\n---\n" + o.stderr[-2000:])
                    print(f"FAIL iter {it}.{sub}: {' '.join(flags)}  (rc ref {r.returncode} ours {o.returncode}) -> {keep}", flush=True)
        if it % 10 == 9:
            print(f"iter {it + 1}: {nfail} failures so far", flush=True)
    print(f"done: {a.iters} iterations, {nfail} failures")
    sys.exit(1 if nfail else 0)


if __name__ == "__main__":
    main()
