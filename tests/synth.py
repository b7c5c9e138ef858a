"""Deterministic synthetic genomes / reads for the tests and the bench (numpy PCG64 streams)."""
from __future__ import annotations

import subprocess
from pathlib import Path

import numpy as np

from helpers import REF_BUILD, REF_DIR, ReadBatch, finalize_batch

CACHE = REF_DIR / "cache"


def synth_genome_survey(n_seqs: int, total_len: int, seed: int) -> list[tuple[str, bytes]]:
    """Benchmark genome in the style of SURVEY.md §8(d) config 3: i.i.d. ACGT with GC 41 %, ~10 % of the bases covered
    by two repeat families (300-bp and 6-kbp elements) whose copies diverge 1-15 % from the consensus, and a few N gaps."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    p = [0.295, 0.205, 0.205, 0.295]
    fams = [rng.choice(acgt, size=300, p=p), rng.choice(acgt, size=6000, p=p)]
    sizes = rng.dirichlet(np.ones(n_seqs) * 3) * total_len
    out = []
    for s in range(n_seqs):
        L = max(20000, int(sizes[s]))
        if total_len >= 1_000_000_000:
            # hg19 scale: a 256-entry table realises the base frequencies (76/52/52/76 of 256) ~30x faster than choice(p=...);
            # smaller genomes keep the original stream so that cached indexes still match their genomes
            lut = np.repeat(acgt, [76, 52, 52, 76])
            g = lut[rng.integers(0, 256, size=L, dtype=np.uint8)]
        else:
            g = rng.choice(acgt, size=L, p=p).copy()
        for fam, cover in ((fams[0], 0.07), (fams[1], 0.03)):
            fl = len(fam)
            for _ in range(max(1, int(L * cover / fl))):
                pos = int(rng.integers(0, L - fl))
                cp = fam.copy()
                mut = rng.random(fl) < rng.uniform(0.01, 0.15)
                cp[mut] = rng.choice(acgt, size=int(mut.sum()))
                g[pos:pos + fl] = cp
        for _ in range(3):
            pos = int(rng.integers(1000, L - 2000))
            g[pos:pos + int(rng.integers(10, 500))] = ord("N")
        out.append((f"chr{s + 1} synthetic len={L}", g.tobytes()))
    return out


def synth_genome(n_seqs: int, total_len: int, seed: int, with_gaps: bool = False, repeats: bool = True) -> list[tuple[str, bytes]]:
    rng = np.random.default_rng(seed)
    sizes = rng.dirichlet(np.ones(n_seqs) * 3) * total_len
    out = []
    fam = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=300, p=[0.295, 0.205, 0.205, 0.295])
    for s in range(n_seqs):
        L = max(2000, int(sizes[s]))
        g = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=L, p=[0.295, 0.205, 0.205, 0.295]).copy()
        if repeats:
            for _ in range(max(1, L // 5000)):
                p = int(rng.integers(0, L - 300))
                cp = fam.copy()
                mut = rng.random(300) < 0.03
                cp[mut] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(mut.sum()))
                g[p:p + 300] = cp
        if with_gaps:
            for _ in range(2):
                p = int(rng.integers(100, L - 200))
                g[p:p + int(rng.integers(1, 60))] = ord("N")
        out.append((f"seq{s} synthetic len={L}", g.tobytes()))
    return out


def write_fasta(path: Path, seqs: list[tuple[str, bytes]]) -> None:
    with open(path, "wb") as f:
        for name, s in seqs:
            f.write(b">" + name.encode() + b"\n")
            if len(s) > 50_000_000:                      # hg19-scale sequences: megabase lines, not 50 M Python iterations
                for i in range(0, len(s), 1 << 20):
                    f.write(s[i:i + (1 << 20)] + b"\n")
                continue
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + b"\n")


def build_synth_index(tag: str, n_seqs: int, total_len: int, seed: int, ftab_chars: int = 10, off_rate: int = 5,
                      with_gaps: bool = False, threads: int = 8, style: str = "flat", builder: str = "reference"):
    """Build (once; cached under oracle/_ref/cache) an index with the reference's bowtie-build, or — builder="gpu" — with
    bt_index_build (same files, minutes instead of hours at hg19 scale; needs a GPU).  Returns (basename, genome)."""
    CACHE.mkdir(parents=True, exist_ok=True)
    base = CACHE / f"{tag}_{n_seqs}_{total_len}_{seed}_{ftab_chars}_{off_rate}_{int(with_gaps)}"
    genome = synth_genome_survey(n_seqs, total_len, seed) if style == "survey" else synth_genome(n_seqs, total_len, seed, with_gaps)
    if not Path(str(base) + ".rev.2.ebwt").exists():
        fa = Path(str(base) + ".fa")
        write_fasta(fa, genome)
        if builder == "gpu":
            import bowtie_b200
            bowtie_b200.build_index(fa, base, off_rate=off_rate, ftab_chars=ftab_chars)
        else:
            p = subprocess.run([str(REF_BUILD), "-q", "-t", str(ftab_chars), "-o", str(off_rate), "--threads", str(threads), str(fa), str(base)],
                               capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError("bowtie-build failed: " + p.stderr)
    return base, genome


_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def synth_reads(genome: list[tuple[str, bytes]], n: int, length: int | tuple[int, int], seed: int, sub_rate: float = 0.01,
                n_rate: float = 0.0, random_frac: float = 0.02, qual_profile: str = "mixed", global_seed: int = 0) -> ReadBatch:
    """Reads sampled uniformly from the genome (both strands) with substitutions, optional Ns and
    a fraction of random (unalignable) reads.  qual_profile: 'high' | 'mixed' | 'low'."""
    rng = np.random.default_rng(seed)
    seqs = [np.frombuffer(s, np.uint8) for _, s in genome]
    w = np.array([len(s) for s in seqs], float)
    w /= w.sum()
    acgt = np.frombuffer(b"ACGT", np.uint8)
    if qual_profile == "high":
        qchoices, qp = np.array([40, 40, 40, 35, 30]), None
    elif qual_profile == "low":
        qchoices, qp = np.array([40, 30, 20, 12, 8, 3, 2]), None
    else:
        qchoices, qp = np.array([40, 40, 40, 35, 30, 20, 10]), None
    names, rs, qs = [], [], []
    for i in range(n):
        L = length if isinstance(length, int) else int(rng.integers(length[0], length[1] + 1))
        if rng.random() < random_frac:
            r = rng.choice(acgt, size=L).copy()
        else:
            while True:
                si = int(rng.choice(len(seqs), p=w))
                if len(seqs[si]) > L:
                    break
            p = int(rng.integers(0, len(seqs[si]) - L))
            r = seqs[si][p:p + L].copy()
            mut = rng.random(L) < sub_rate
            r[mut] = rng.choice(acgt, size=int(mut.sum()))
            if rng.random() < 0.5:
                r = np.frombuffer(r.tobytes().translate(_COMP)[::-1], np.uint8).copy()
        if n_rate > 0:
            r[rng.random(L) < n_rate] = ord("N")
        q = (rng.choice(qchoices, size=L, p=qp) + 33).astype(np.uint8)
        names.append(f"s{i}".encode())
        rs.append(r.tobytes())
        qs.append(q.tobytes())
    return finalize_batch(names, rs, qs, global_seed)


def write_fastq(path: Path, batch: ReadBatch) -> None:
    with open(path, "wb") as f:
        for n, s, q in zip(batch.names, batch.seqs, batch.quals):
            f.write(b"@" + n + b"\n" + s + b"\n+\n" + q + b"\n")


def synth_pairs(genome: list[tuple[str, bytes]], n: int, length: int | tuple[int, int], seed: int, frag_mean: float = 200.0, frag_sd: float = 20.0,
                sub_rate: float = 0.02, n_rate: float = 0.0, qual_profile: str = "mixed", broken_frac: float = 0.05):
    """Paired reads in --fr orientation: mate 1 from the fragment's left end (forward), mate 2 the reverse complement of
    its right end; the whole fragment is flipped half of the time.  `broken_frac` of the pairs get an unrelated mate 2.
    Returns (names, seqs1, quals1, seqs2, quals2) as lists of bytes."""
    rng = np.random.default_rng(seed)
    seqs = [np.frombuffer(s, np.uint8) for _, s in genome]
    w = np.array([len(s) for s in seqs], float)
    w /= w.sum()
    acgt = np.frombuffer(b"ACGT", np.uint8)
    qchoices = {"high": [40, 40, 40, 35, 30], "low": [40, 30, 20, 12, 8, 3, 2]}.get(qual_profile, [40, 40, 40, 35, 30, 20, 10])

    def mutate(r):
        r = r.copy()
        mut = rng.random(len(r)) < sub_rate
        r[mut] = rng.choice(acgt, size=int(mut.sum()))
        if n_rate > 0:
            r[rng.random(len(r)) < n_rate] = ord("N")
        return r

    def rc(r):
        return np.frombuffer(r.tobytes().translate(_COMP)[::-1], np.uint8)

    names, s1, q1, s2, q2 = [], [], [], [], []
    for i in range(n):
        L1 = length if isinstance(length, int) else int(rng.integers(length[0], length[1] + 1))
        L2 = length if isinstance(length, int) else int(rng.integers(length[0], length[1] + 1))
        while True:
            si = int(rng.choice(len(seqs), p=w))
            F = max(max(L1, L2) + 1, int(rng.normal(frag_mean, frag_sd)))
            if len(seqs[si]) > F + 1:
                break
        p = int(rng.integers(0, len(seqs[si]) - F))
        frag = seqs[si][p:p + F]
        if rng.random() < 0.5:
            frag = rc(frag)
        a, b = mutate(frag[:L1]), mutate(rc(frag[F - L2:]))
        if rng.random() < broken_frac:
            b = rng.choice(acgt, size=L2)
        names.append(f"p{i}".encode())
        s1.append(a.tobytes()); s2.append(b.tobytes())
        q1.append((rng.choice(qchoices, size=L1) + 33).astype(np.uint8).tobytes())
        q2.append((rng.choice(qchoices, size=L2) + 33).astype(np.uint8).tobytes())
    return names, s1, q1, s2, q2


def write_fastq_pairs(path1: Path, path2: Path, pairs) -> None:
    names, s1, q1, s2, q2 = pairs
    with open(path1, "wb") as f1, open(path2, "wb") as f2:
        for n, a, qa, b, qb in zip(names, s1, q1, s2, q2):
            f1.write(b"@" + n + b"/1\n" + a + b"\n+\n" + qa + b"\n")
            f2.write(b"@" + n + b"/2\n" + b + b"\n+\n" + qb + b"\n")
