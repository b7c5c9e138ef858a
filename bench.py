#!/usr/bin/env python
"""bench.py — aligned reads/s of the B200 FM-index backward-search path (100 bp, -n 2) on an hg19-sized index.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
    python bench.py --impl reference [--gpus N] ...               the reference's CPU path (oracle/_ref)

Workload (BASELINE.json configs 3-5): a synthetic 3.0-Gbp genome (24 sequences sized like the hg19 chromosomes,
GC 41 %, 10 % of the bases in 300-bp / 6-kbp repeat families with 1-15 % divergence, N gaps), indexed ONCE per box by
bt_index_build_text (the GPU index builder; files byte-identical to bowtie-build's, tests/test_index_build.py) into
/dev/shm — whichever arm runs first builds, the other reuses — and 100-bp reads sampled from it (1 % substitutions, both
strands, 0.1 % random).  Headline: `-n 2 -k 1` (config 4).  `other_policies` carries `-n 2 --best` (config 3) and paired
`-n 3` (config 5) measured the same way after the headline.

A "step" is one pass of the hot path (bt_align_batch*, all phases of the policy) over one batch of reads.  `value` = whole-job
reads/s with the batch already resident in HBM; `e2e` = the same through the C ABI with pinned HOST buffers (H2D + D2H inside
the timed region).  `roofline` is the search kernel's algorithmic bytes (64 B per side fetch + 4 B per offs[]
read + 8 B per ftab jump, counted by the kernel itself) over its device time.  `cpu_baseline` / the reference arm: the
unmodified reference binary on the host cores this process may use, best of a -p sweep, on a search-time basis (wall clock
minus the wall clock of a one-read run = index load), with the wall-clock figure beside it.  `parity_sample`: the reference's
output for the CPU sample compared field by field with the GPU's records for the same reads; a mismatch fails the run.
Under torchrun every rank owns one GPU, holds the whole index and a disjoint shard of reads (weak scaling); the only
collective is the all-reduce of the five hit counters.
"""
from __future__ import annotations

import argparse
import fcntl
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
REF_DIR = ROOT / "oracle" / "_ref"
READ_LEN = 100
HG19_CHR = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022, 141213431, 135534747, 135006516,
            133851895, 115169878, 107349540, 102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566,
            155270560, 59373566]


# ------------------------------------------------------------------------------------------------
# the hg19-sized synthetic index
# ------------------------------------------------------------------------------------------------

def hg19_like_genome(total_len: int, n_seqs: int = 24, seed: int = 1):
    """SURVEY.md §8(d) config 3 genome, vectorised: returns (codes uint8[N] of the joined unambiguous text, records
    [(off, len, first)], names).  i.i.d. bases with GC 40.6 % (a 256-entry table: 76/52/52/76), 7 % of the bases covered by
    eight 300-bp repeat families and 3 % by two 6-kbp families (copies 1-15 % diverged from their consensus, <= 10^5 copies
    per family at 3 Gbp), three N gaps of 10-500 characters per sequence (so nFrag > nPat)."""
    rng = np.random.default_rng(seed)
    w = np.array([HG19_CHR[i % 24] for i in range(n_seqs)], float)
    sizes = np.maximum(20000, (w / w.sum() * total_len).astype(np.int64))
    N = int(sizes.sum())
    lut = np.repeat(np.arange(4, dtype=np.uint8), [76, 52, 52, 76])
    g = np.empty(N, np.uint8)
    CH = 1 << 27
    for a in range(0, N, CH):
        b = min(N, a + CH)
        g[a:b] = lut[rng.integers(0, 256, size=b - a, dtype=np.uint8)]
    for fl, cover, nfam in ((300, 0.07, 8), (6000, 0.03, 2)):
        fams = lut[rng.integers(0, 256, size=(nfam, fl), dtype=np.uint8)]
        c = max(nfam, int(N * cover / fl))
        slot = N // c                                           # one copy per slot, jittered inside it: copies of a length never overlap
        pos = np.arange(c, dtype=np.int64) * slot + rng.integers(0, max(1, slot - fl + 1), size=c)
        pos = np.minimum(pos, N - fl)
        fam_of = rng.integers(0, nfam, size=c)
        div = rng.uniform(0.01, 0.15, size=c).astype(np.float32)
        step = max(1, (1 << 24) // fl)
        ar = np.arange(fl, dtype=np.int64)
        for a in range(0, c, step):
            b = min(c, a + step)
            cp = fams[fam_of[a:b]].copy()
            mut = rng.random((b - a, fl), dtype=np.float32) < div[a:b, None]
            cp[mut] = rng.integers(0, 4, size=int(mut.sum()), dtype=np.uint8)
            g[pos[a:b, None] + ar[None, :]] = cp
    recs, names = [], []
    for s in range(n_seqs):
        L = int(sizes[s])
        cuts = np.sort(rng.choice(np.arange(1000, L - 1000), size=3, replace=False))
        gaps = rng.integers(10, 500, size=3)
        recs.append((0, int(cuts[0]), 1))
        recs.append((int(gaps[0]), int(cuts[1] - cuts[0]), 0))
        recs.append((int(gaps[1]), int(cuts[2] - cuts[1]), 0))
        recs.append((int(gaps[2]), int(L - cuts[2]), 0))
        names.append(f"chr{s + 1}")
    return g, recs, names


def cache_dir() -> Path:
    env = os.environ.get("BT_BENCH_CACHE")
    if env:
        return Path(env)
    try:
        if shutil.disk_usage("/dev/shm").free > 12 << 30:
            return Path("/dev/shm/bowtie_b200_bench")
    except Exception:
        pass
    return Path(tempfile.gettempdir()) / "bowtie_b200_bench"


def make_index(base: Path, mbp: int, device: int) -> None:
    """--make-index: generate the genome and build the six index files (runs in its own process so that neither arm's
    measuring process holds the builder's memory — and the reference arm never loads the product library)."""
    import bowtie_b200
    t0 = time.time()
    g, recs, names = hg19_like_genome(mbp * 1_000_000, 24, 1)
    t1 = time.time()
    if not os.environ.get("BOWTIE_B200_LIB"):
        bowtie_b200.build_library()
    bowtie_b200.build_index_text(g, recs, names, base, off_rate=5, ftab_chars=10, device=device)
    t2 = time.time()
    Path(str(base) + ".json").write_text(json.dumps({"genome_s": round(t1 - t0, 1), "build_s": round(t2 - t1, 1), "len": int(g.size),
                                                     "builder": "bt_index_build_text (B200)"}))


def ensure_index(mbp: int, device: int) -> tuple[Path, str, dict]:
    """The bench index: BT_BENCH_INDEX if given; else the hg19-sized index under the cache directory, built once per box
    (file lock: under torchrun or with both arms on one box, one process builds and the others wait)."""
    env = os.environ.get("BT_BENCH_INDEX")
    if env:
        return Path(env), Path(env).name, {}
    d = cache_dir()
    d.mkdir(parents=True, exist_ok=True)
    base = d / f"hg19s_{mbp}m_24_1_10_5"
    done = Path(str(base) + ".done")
    info: dict = {}
    with open(d / "build.lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        if not done.exists():
            t0 = time.time()
            p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--make-index", str(base), "--mbp", str(mbp), "--device", str(device)],
                               capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError("building the bench index failed: " + (p.stderr or p.stdout)[-800:])
            done.write_text(f"{time.time() - t0:.1f}\n")
            info["built_now_s"] = round(time.time() - t0, 1)
    try:
        info.update(json.loads(Path(str(base) + ".json").read_text()))
    except Exception:
        pass
    return base, base.name, info


def fallback_index() -> tuple[Path, str]:
    """Largest prebuilt synthetic index under oracle/_ref/cache (built with the reference's bowtie-build), else the shipped e_coli."""
    best = None
    for p in sorted((REF_DIR / "cache").glob("bench*.rev.2.ebwt")):
        base = Path(str(p)[: -len(".rev.2.ebwt")])
        if Path(str(base) + ".4.ebwt").exists():
            sz = Path(str(base) + ".1.ebwt").stat().st_size
            if best is None or sz > best[0]:
                best = (sz, base)
    if best:
        return best[1], best[1].name
    base = REF_DIR / "fixtures" / "e_coli"
    if Path(str(base) + ".1.ebwt").exists():
        return base, "e_coli (reference fixture)"
    raise SystemExit("bench.py: no index available")


def load_genome(base: Path) -> np.ndarray:
    """Base codes 0..3 of the joined reference from X.4.ebwt (2-bit packed, base i in bits 2*(i&3) of byte i>>2;
    reference.h:58-330).  Reads are sampled from this text, so no FASTA has to travel with the index."""
    raw = np.fromfile(str(base) + ".4.ebwt", np.uint8)
    out = np.empty(len(raw) * 4, np.uint8)
    for k in range(4):
        out[k::4] = (raw >> (2 * k)) & 3
    return out


def index_len(base: Path) -> int:
    with open(str(base) + ".1.ebwt", "rb") as f:
        return int(np.frombuffer(f.read(8), np.uint32)[1])


# ------------------------------------------------------------------------------------------------
# reads
# ------------------------------------------------------------------------------------------------

def _seeds(codes2d, quals2d, name2d):
    """Read::seed per pat.cpp:21-57 with global seed 0, vectorised."""
    n, L = codes2d.shape
    seeds = np.full(n, ((0 + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83) & 0xFFFFFFFF, np.uint32)

    def fold(a, period, shift):
        """XOR over i of a[:, i] << (shift * (i % period)): columns with the same shift are XORed as bytes first."""
        w = a.shape[1]
        acc = np.zeros((n, period), np.uint8)
        for k in range(0, w, period):
            blk = a[:, k:k + period]
            acc[:, :blk.shape[1]] ^= blk
        return np.bitwise_xor.reduce(acc.astype(np.uint32) << (np.arange(period, dtype=np.uint32) * np.uint32(shift)), axis=1)
    seeds ^= fold(codes2d, 16, 2)
    seeds ^= fold(quals2d, 4, 8)
    seeds ^= fold(name2d, 4, 8)
    return seeds.astype(np.uint32)


_QUAL_LUT = (np.array([40, 40, 40, 35, 30, 20, 10], np.uint8)[(np.arange(256) * 7) >> 8] + 33).astype(np.uint8)   # random byte -> Phred+33 character
GEN_CHUNK = 1 << 20            # reads generated per numpy pass: bounds the temporaries (a (n, 100) int64 index array) whatever the batch size


def _chunked(gen, genome: np.ndarray, n: int, seed: int, per_unit: int):
    """Runs a generator over [0, n) in chunks of GEN_CHUNK units (chunk k: seed (seed, k), read ids continuing) and concatenates."""
    if n <= GEN_CHUNK:
        return gen(genome, n, seed, 0)
    parts = [gen(genome, min(GEN_CHUNK, n - a), (seed, a // GEN_CHUNK), a) for a in range(0, n, GEN_CHUNK)]
    offs = (np.arange(per_unit * n + 1, dtype=np.uint64) * READ_LEN)
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), offs, np.concatenate([p[3] for p in parts]),
            np.concatenate([p[4] for p in parts]))


def make_reads(genome: np.ndarray, n: int, seed: int):
    """n x 100 bp: uniform positions, both strands, 1 % substitutions, 0.1 % random reads, Phred33 quals from
    {40,40,40,35,30,20,10}.  Returns (codes[n*100], quals[n*100], offs[n+1], seeds[n], names)."""
    return _chunked(_make_reads, genome, n, seed, 1)


def make_pairs(genome: np.ndarray, n: int, seed: int):
    """n pairs of 2 x 100 bp in --fr geometry: fragment length N(200, 20) clipped to [101, 250], mate 1 the
    fragment's first 100 bases, mate 2 the reverse complement of its last 100, the whole fragment flipped half of the time,
    1 % substitutions.  Reads are interleaved (mate 1, mate 2, mate 1, ...); names "r%09d/1", "r%09d/2"."""
    return _chunked(_make_pairs, genome, n, seed, 2)


def _make_reads(genome: np.ndarray, n: int, seed, id0: int):
    rng = np.random.default_rng(seed)
    L = READ_LEN
    pos = rng.integers(0, len(genome) - L, n)
    codes = np.lib.stride_tricks.sliding_window_view(genome, L)[pos]        # (n, L) gather of windows, no index matrix
    mi = rng.integers(0, n * L, int(rng.binomial(n * L, 0.01)))              # 1 % substitutions (positions drawn directly)
    cf = codes.reshape(-1)
    cf[mi] = (cf[mi] + rng.integers(1, 4, mi.size, dtype=np.uint8)) & 3
    rnd = rng.random(n) < 0.001
    codes[rnd] = rng.integers(0, 4, (int(rnd.sum()), L), dtype=np.uint8)
    rc = rng.random(n) < 0.5
    flipped = codes[rc, ::-1]
    codes[rc] = np.where(flipped < 4, 3 - flipped, 4)
    codes = np.ascontiguousarray(codes, np.uint8)
    quals = _QUAL_LUT[np.frombuffer(rng.bytes(n * L), np.uint8)].reshape(n, L)
    ids = np.arange(id0, id0 + n, dtype=np.uint32)
    name = np.zeros((n, 10), np.uint8)                           # "r%09d": fixed width so genRandSeed vectorises
    name[:, 0] = ord("r")
    for d in range(9):
        name[:, 9 - d] = (ids // 10 ** d) % 10 + 48
    offs = (np.arange(n + 1, dtype=np.uint64) * L)
    return codes.reshape(-1), quals.reshape(-1), offs, _seeds(codes, quals, name), name


def _make_pairs(genome: np.ndarray, n: int, seed, id0: int):
    rng = np.random.default_rng(seed)
    L = READ_LEN
    F = np.clip(rng.normal(200.0, 20.0, n).round().astype(np.int64), L + 1, 250)
    pos = rng.integers(0, len(genome) - 251, n)
    win = np.lib.stride_tricks.sliding_window_view(genome, L)
    left = win[pos]
    right = win[pos + F - L]
    rcomp = lambda a: np.where(a[:, ::-1] < 4, 3 - a[:, ::-1], 4)
    flip = rng.random(n) < 0.5
    m1 = np.where(flip[:, None], rcomp(right), left)
    m2 = np.where(flip[:, None], left, rcomp(right))
    codes = np.empty((2 * n, L), np.uint8)
    codes[0::2] = m1; codes[1::2] = m2
    mi = rng.integers(0, 2 * n * L, int(rng.binomial(2 * n * L, 0.01)))
    cf = codes.reshape(-1)
    cf[mi] = (cf[mi] + rng.integers(1, 4, mi.size, dtype=np.uint8)) & 3
    quals = _QUAL_LUT[np.frombuffer(rng.bytes(2 * n * L), np.uint8)].reshape(2 * n, L)
    ids = np.repeat(np.arange(id0, id0 + n, dtype=np.uint32), 2)
    name = np.zeros((2 * n, 12), np.uint8)
    name[:, 0] = ord("r")
    for d in range(9):
        name[:, 9 - d] = (ids // 10 ** d) % 10 + 48
    name[:, 10] = ord("/"); name[0::2, 11] = ord("1"); name[1::2, 11] = ord("2")
    offs = (np.arange(2 * n + 1, dtype=np.uint64) * L)
    return codes.reshape(-1), quals.reshape(-1), offs, _seeds(codes, quals, name), name


def _fastq_records(codes, quals, name, nreads: int) -> np.ndarray:
    L = READ_LEN
    lut = np.frombuffer(b"ACGTN", np.uint8)
    seq = lut[codes[: nreads * L]].reshape(nreads, L)
    q = quals[: nreads * L].reshape(nreads, L)
    W = name.shape[1]
    rec = np.empty((nreads, 1 + W + 1 + L + 3 + L + 1), np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1:1 + W] = name[:nreads]; rec[:, 1 + W] = 10
    rec[:, 2 + W:2 + W + L] = seq; rec[:, 2 + W + L] = 10; rec[:, 3 + W + L] = ord("+"); rec[:, 4 + W + L] = 10
    rec[:, 5 + W + L:5 + W + 2 * L] = q; rec[:, 5 + W + 2 * L] = 10
    return rec


def write_sample(td: Path, h, n: int, paired: bool):
    """FASTQ file(s) of the first n reads / pairs of a batch."""
    if paired:
        rec = _fastq_records(h[0], h[1], h[4], 2 * n)
        (td / "s_1.fq").write_bytes(rec[0::2].tobytes()); (td / "s_2.fq").write_bytes(rec[1::2].tobytes())
        return (td / "s_1.fq", td / "s_2.fq")
    (td / "s.fq").write_bytes(_fastq_records(h[0], h[1], h[4], n).tobytes())
    return td / "s.fq"


# ------------------------------------------------------------------------------------------------
# clocks (recipe of /opt/skills/guides/B200_PROFILING.md), sampled through NVML by rank 0 only
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    def __init__(self, gpus: list[int], enabled: bool = True) -> None:
        self.gpus, self.samples, self.stop, self.enabled = gpus, [], threading.Event(), enabled
        self.t = threading.Thread(target=self.run, daemon=True)
        self.nv = None
        self.max_mhz = None

    def run(self) -> None:
        try:
            import pynvml as nv
            nv.nvmlInit()
            hs = [nv.nvmlDeviceGetHandleByIndex(g) for g in self.gpus]
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(hs[0], nv.NVML_CLOCK_SM))
            while not self.stop.is_set():
                for h in hs:
                    try:
                        self.samples.append((int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))))
                    except Exception:
                        pass
                self.stop.wait(0.2)
        except Exception:
            self.run_smi()

    def run_smi(self) -> None:
        Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        bits = [0x8, 0x40, 0x20, 0x4]
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpus[0]), f"--query-gpu={Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                mask = sum(b for b, s in zip(bits, o[2:6]) if s.strip().startswith("Active"))
                self.samples.append((int(o[0]), mask)); self.max_mhz = int(o[1])
            except Exception:
                pass
            self.stop.wait(0.5)

    def __enter__(self):
        if self.enabled:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.enabled:
            self.t.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(s[0] for s in self.samples)
        allmask = 0
        for s in self.samples:
            allmask |= s[1]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": [n for b, n in names.items() if allmask & b],
                "samples": len(self.samples), "gpus_sampled": len(self.gpus)}


# ------------------------------------------------------------------------------------------------
# the reference on the host cores (reference arm / cpu_baseline) and the parity check
# ------------------------------------------------------------------------------------------------

def host_cores() -> int:
    """Cores this process may actually use: the affinity mask, capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        try:
            q = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text()); p = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                n = min(n, max(1, q // p))
        except Exception:
            pass
    return max(1, n)


def cpu_model() -> str:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(base: Path, fq, threads: int, flags, out_path="/dev/null", stderr_to: list | None = None) -> tuple[float, float]:
    """bowtie-align-s <flags> -t -p <threads>; returns (wall seconds, 'Time searching' seconds or -1); the run's stderr is
    appended to `stderr_to` when given."""
    exe = REF_DIR / "bowtie-align-s"
    if not exe.exists():
        raise RuntimeError("oracle/_ref/bowtie-align-s missing (built by oracle/Makefile from /root/reference)")
    inputs = ["-1", str(fq[0]), "-2", str(fq[1])] if isinstance(fq, (tuple, list)) else [str(fq)]
    t0 = time.time()
    p = subprocess.run([str(exe), *flags, "-t", "-p", str(threads), "-x", str(base), *inputs, str(out_path)], capture_output=True, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("reference run failed: " + p.stderr[-500:])
    if stderr_to is not None:
        stderr_to.append(p.stderr)
    search = -1.0
    for line in (p.stdout + p.stderr).splitlines():
        if line.startswith("Time searching:"):
            h, m, s = line.split(":", 1)[1].strip().split(":")
            search = float(int(h) * 3600 + int(m) * 60 + int(s))
    return wall, search


class ReferenceRunner:
    """The reference's own pthreads path on this box: measures the fixed cost of a run (process start + index load: a one-read
    run), picks -p from {cores/4, cores/2, cores} on a small sample, then times samples on a search-time basis."""

    def __init__(self, base: Path, flags, paired: bool, td: Path, h, rate_hint: float | None = None) -> None:
        self.base, self.flags, self.paired, self.td, self.h = base, list(flags), paired, td, h
        self.cores = host_cores()
        (td / "one").mkdir(exist_ok=True)
        one = write_sample(td / "one", h, 1, paired)
        self.overhead = min(run_reference(base, one, 1, self.flags)[0] for _ in range(2))
        self.sweep: dict[int, float] = {}
        self.threads = self.cores

    def pick_threads(self, n_sweep: int) -> None:
        (self.td / "sw").mkdir(exist_ok=True)
        fq = write_sample(self.td / "sw", self.h, n_sweep, self.paired)
        cands = sorted({max(1, self.cores // 4), max(1, self.cores // 2), self.cores})
        for p in cands:
            wall, _ = run_reference(self.base, fq, p, self.flags)
            self.sweep[p] = n_sweep / max(1e-3, wall - self.overhead)
        self.threads = max(self.sweep, key=self.sweep.get)

    def time_sample(self, fq, n: int, out_path="/dev/null") -> dict:
        err: list = []
        wall, search = run_reference(self.base, fq, self.threads, self.flags, out_path, stderr_to=err)
        import re
        gave_up = {int(m) for m in re.findall(r"Exhausted best-first chunk memory for read \S+ \(patid (\d+)\)", err[0])}
        return {"n": n, "wall_s": wall, "search_s": max(1e-3, wall - self.overhead), "time_searching_s": search, "gave_up": gave_up}


def cli_e2e(base: Path, h, n: int, td: Path, rr: "ReferenceRunner") -> dict:
    """FASTQ file in -> SAM file out through the drop-in program (`bowtie-b200-align`: device read ingest, search, device formatting)
    against `bowtie-align-s -p <best>` on the same file; both on a search-time basis (wall clock minus a one-read run of the same
    program = process start + index load), with the wall-clock figures beside them.  Files live in RAM-backed storage."""
    exe = ROOT / "bowtie_b200" / "bowtie-b200-align"
    (td / "cli").mkdir(exist_ok=True)
    fq = write_sample(td / "cli", h, n, False)
    one = td / "one" / "s.fq"
    flags = ["-n", "2", "-k", "1", "-S"]

    own: dict = {}

    def run(cmd, timing: bool = False) -> float:
        t0 = time.time()
        p = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, BT_CLI_TIMING="1") if timing else None)
        if p.returncode != 0:
            raise RuntimeError("cli_e2e run failed: " + p.stderr[-300:])
        if timing:                     # the program's own clock: "timing: index load X s, reads to output Y s"
            m = re.search(r"timing: index load ([0-9.]+) s, reads to output ([0-9.]+) s", p.stderr)
            if m:
                own["index_load_s"], own["reads_to_output_s"] = float(m.group(1)), float(m.group(2))
            m = re.search(r"device I/O chunks: [^\n]*", p.stderr)
            if m:
                own["device_io"] = m.group(0)
        return time.time() - t0
    o1 = min(run([str(exe), *flags, "-x", str(base), str(one), str(td / "o1.sam")]) for _ in range(2))
    w = run([str(exe), *flags, "-x", str(base), str(fq), str(td / "ours.sam")], timing=True)
    r1 = min(run([str(REF_DIR / "bowtie-align-s"), *flags, "-p", str(rr.threads), "-x", str(base), str(one), str(td / "r1.sam")]) for _ in range(2))
    rw = run([str(REF_DIR / "bowtie-align-s"), *flags, "-p", str(rr.threads), "-x", str(base), str(fq), str(td / "ref.sam")])
    same = None
    try:
        a = [l for l in (td / "ours.sam").read_bytes().split(b"\n") if not l.startswith(b"@PG")]
        if rr.threads == 1:
            b = [l for l in (td / "ref.sam").read_bytes().split(b"\n") if not l.startswith(b"@PG")]
            same = a == b
        else:                      # the reference's -p threads write in completion order: compare as multisets of lines
            b = [l for l in (td / "ref.sam").read_bytes().split(b"\n") if not l.startswith(b"@PG")]
            same = sorted(a) == sorted(b)
    except Exception:
        pass
    for f in ("ours.sam", "ref.sam"):
        try:
            (td / f).unlink()
        except Exception:
            pass
    if "reads_to_output_s" in own:
        own["value_by_own_clock"] = n / max(1e-3, own["reads_to_output_s"])       # index load varies by seconds between two runs on one box: the program's own split is the steadier basis
    return {"reads": n, "flags": " ".join(flags), "value": n / max(1e-3, w - o1), "unit": "reads/s", "wall_s": round(w, 2), "start_and_index_load_s": round(o1, 2), "own_clock": own,
            "reference_value": n / max(1e-3, rw - r1), "reference_wall_s": round(rw, 2), "reference_start_and_index_load_s": round(r1, 2),
            "reference_threads": rr.threads, "sam_identical": same}


def parse_reference_output(path: Path, refnames: list[str]) -> dict:
    """Default-format hit lines (hit.cpp:176-240) -> {(unit, mate): (fw, tidx, toff, oms, ((pos, refc), ...))}."""
    tid = {n.split()[0] if n.split() else n: i for i, n in enumerate(refnames)}
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    out = {}
    with open(path) as f:
        for line in f:
            p = line.rstrip("\n").split("\t")
            name = p[0]
            mate = 0
            if len(name) > 10 and name[10] == "/":
                mate = int(name[11]); name = name[:10]
            mms = ()
            if len(p) > 7 and p[7]:
                mms = tuple(sorted((int(x.split(":")[0]), code[x.split(":")[1][0]]) for x in p[7].split(",")))
            out[(int(name[1:]), mate)] = (p[1] == "+", tid[p[2]], int(p[3]), int(p[6]), mms)
    return out


def gpu_records(found, hits, n_units: int, paired: bool) -> dict:
    """The same dictionary from the library's records (include/bowtie_b200.h: hit record layout) for -k 1."""
    out = {}
    slots = hits.shape[1]
    for u in np.nonzero(found[:n_units] > 0)[0]:
        for s in range(min(int(found[u]), slots)):
            w = hits[u, s]
            mate = int((w[3] >> 25) & 3)
            nmm = int(w[4])
            mms = tuple(sorted((int(x & 0xffff), int((x >> 16) & 0xff)) for x in w[5:5 + nmm]))
            out[(int(u), mate)] = (bool((w[3] >> 24) & 1), int(w[0]), int(w[1]), int(w[2]), mms)
    return out


def compare_parity(ref: dict, got: dict, gave_up=frozenset()) -> dict:
    """`gave_up`: units for which the reference printed "Exhausted best-first chunk memory ... skipping read" (pool.h:146-165) — it
    dropped the rest of their search, the library finishes it; they are counted apart and do not fail the run."""
    bad = [k for k in set(ref) | set(got) if ref.get(k) != got.get(k)]
    excused = [k for k in bad if k[0] in gave_up]
    bad = [k for k in bad if k[0] not in gave_up]
    return {"records_reference": len(ref), "records_gpu": len(got), "mismatching": len(bad),
            "reference_gave_up_units": len(gave_up), "records_differing_in_those": len(excused),
            **({"first_mismatch": str((bad[0], ref.get(bad[0]), got.get(bad[0])))} if bad else {})}


# ------------------------------------------------------------------------------------------------

POLICIES = {
    "n2k1":   dict(flags=["-n", "2", "-k", "1"], paired=False, metric="aligned reads/sec (100 bp, -n 2)", cfg="config 4"),
    "best":   dict(flags=["-n", "2", "--best"], paired=False, metric="aligned reads/sec (100 bp, -n 2 --best)", cfg="config 3"),
    "paired": dict(flags=["-n", "3"], paired=True, metric="aligned read pairs/sec (2x100 bp, -n 3 paired-end)", cfg="config 5"),
    "v0":     dict(flags=["-v", "0"], paired=False, metric="aligned reads/sec (100 bp, -v 0)", cfg="config 2"),
}


def lib_policy(name: str):
    import bowtie_b200
    if name == "v0":
        return bowtie_b200.Policy(mode=0, mms=0, khits=1)
    return bowtie_b200.Policy(mode=1, mms=3 if name == "paired" else 2, khits=1, best=(name == "best"), paired=(name == "paired"))


def reference_arm(args, base: Path, idx_name: str, idx_info: dict) -> None:
    pd = POLICIES[args.policy]
    R = 2 if pd["paired"] else 1
    unit = "pairs/s" if R == 2 else "reads/s"
    genome = load_genome(base)
    gen = make_pairs if R == 2 else make_reads
    nmax = max(args.cpu_sample, 200_000)
    h0 = gen(genome, nmax, seed=12345)
    del genome
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None) as tdn:
        td = Path(tdn)
        rr = ReferenceRunner(base, pd["flags"], R == 2, td, h0)
        rr.pick_threads(min(nmax, 200_000))
        rate = rr.sweep[rr.threads]
        # a step = one run on a bounded sample: about --cpu-step-seconds of search, at most the prepared reads
        n = int(min(nmax, max(50_000, rate * args.cpu_step_seconds)))
        (td / "st").mkdir()
        fq = write_sample(td / "st", h0, n, R == 2)
        res = []
        for it in range(args.warmup + args.steps):
            r = rr.time_sample(fq, n)
            if it >= args.warmup:
                res.append(r)
    tot_search = sum(r["search_s"] for r in res); tot_wall = sum(r["wall_s"] for r in res)
    ts = [r["time_searching_s"] for r in res if r["time_searching_s"] >= 0]
    val = n * len(res) / tot_search
    cpu = {"value": val, "unit": unit, "cores": rr.threads, "kind": "reference", "host_cores": rr.cores, "cpu_model": cpu_model(),
           "p_sweep": {str(k): round(v) for k, v in rr.sweep.items()},
           "sample": f"{n} {'pairs' if R == 2 else 'reads'} per step, bowtie-align-s {' '.join(pd['flags'])} -p {rr.threads}; search-time basis = wall clock minus a "
                     f"one-read run ({rr.overhead:.2f} s: process start + index load)",
           "value_wall_clock": n * len(res) / tot_wall, "value_time_searching": (n * len(ts) / sum(ts)) if ts and sum(ts) > 0 else None}
    line = {"impl": "reference", "metric": pd["metric"], "value": val, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_search / len(res), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(pd, idx_name), "reads_per_step": n, "index_len_bp": index_len(base), "index": idx_info,
                       "parallelism": f"{rr.threads} of {rr.cores} usable host threads (bowtie -p, best of the sweep)"},
            "cpu_baseline": cpu, "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(pd: dict, idx_name: str) -> str:
    R = 2 if pd["paired"] else 1
    return (f"{' '.join(pd['flags'])}, {'2x' if R == 2 else ''}{READ_LEN} bp synthetic {'pairs (--fr, fragments N(200,20))' if R == 2 else 'reads'} "
            f"(1% subs, both strands), index {idx_name} ({pd['cfg']})")


class Arm:
    """One policy on one rank: resident batches, contexts, the device-timed and the end-to-end loops."""
    batches: dict = {}

    def __init__(self, ix, name: str, genome, B: int, NS: int, rank: int, world: int, local: int) -> None:
        import torch
        import bowtie_b200
        self.torch, self.ix, self.name, self.B, self.rank, self.world = torch, ix, name, B, rank, world
        pd = POLICIES[name]
        self.R = 2 if pd["paired"] else 1
        self.pol = lib_policy(name)
        self.slots, self.mm_cap = self.R, 7
        self.rw = bowtie_b200.BT_HIT_HDR_WORDS + self.mm_cap
        gen = make_pairs if self.R == 2 else make_reads
        # two distinct batches per rank, alternated, each larger than L2 (kept for the strong-scaling pass, which reuses them)
        key = (self.R, B, rank)
        if key not in Arm.batches:
            Arm.batches[key] = [gen(genome, B, seed=12345 + 1000 * rank + k) for k in range(2)]
        self.host = Arm.batches[key]
        self.dev = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (h[0], h[1], h[2].view(np.int64), h[3].view(np.int32))) for h in self.host]
        self.NS = NS
        self.ctxs = [bowtie_b200.Context(ix) for _ in range(NS)]
        self.streams = [torch.cuda.Stream() for _ in range(NS)]
        self.d_out = [(torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda"),
                       torch.zeros(B * self.slots * self.rw, dtype=torch.int32, device="cuda")) for _ in range(NS)]
        self.main = torch.cuda.current_stream()
        self.pin, self.outs = None, None

    def close(self) -> None:
        self.torch.cuda.synchronize()
        for c in self.ctxs:
            c.close()
        self.dev = self.d_out = self.pin = self.outs = None
        self.torch.cuda.empty_cache()

    def step_dev(self, k: int) -> None:
        s, q, o, sd = self.dev[k & 1]
        f, g, h = self.d_out[k % self.NS]
        self.ctxs[k % self.NS].align_device(s.data_ptr(), q.data_ptr(), o.data_ptr(), sd.data_ptr(), self.B * self.R, READ_LEN, self.pol,
                                            f.data_ptr(), g.data_ptr(), h.data_ptr(), self.slots, self.mm_cap, self.streams[k % self.NS].cuda_stream)

    def prepare_e2e(self) -> None:
        torch = self.torch
        self.pin = [[torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (h[0], h[1], h[2].view(np.int64), h[3].view(np.int32))] for h in self.host]
        self.outs = []
        for _ in range(self.NS):
            o = [torch.zeros(self.B, dtype=torch.int32).pin_memory(), torch.zeros(self.B, dtype=torch.int32).pin_memory(),
                 torch.zeros(self.B * self.slots * self.rw, dtype=torch.int32).pin_memory()]
            self.outs.append((o[0].numpy().view(np.uint32), o[1].numpy().view(np.uint32), o[2].numpy().view(np.uint32).reshape(self.B, self.slots, self.rw)))

    def step_e2e(self, k: int) -> None:
        s, q, o, sd = self.pin[k & 1]
        self.ctxs[k % self.NS].align_async(s.numpy(), q.numpy(), o.numpy().view(np.uint64), sd.numpy().view(np.uint32), self.pol, self.outs[k % self.NS],
                                           self.slots, self.mm_cap, self.streams[k % self.NS].cuda_stream)

    def run_steps(self, fn, nsteps: int):
        """nsteps round-robin over NS streams, bracketed by events on the main stream."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.main)
        for st in self.streams:
            st.wait_event(e0)
        for k in range(nsteps):
            fn(k)
        for cx, st in zip(self.ctxs, self.streams):
            cx.join(st.cuda_stream)            # the batch's heavy / overflow passes (and D2H) run on the context's side stream
            ev = torch.cuda.Event()
            ev.record(st)
            self.main.wait_event(ev)
        e1.record(self.main)
        return e0, e1

    @property
    def h2d(self) -> int:
        return 2 * self.B * self.R * READ_LEN + 8 * (self.B * self.R + 1) + 4 * self.B * self.R

    @property
    def d2h(self) -> int:
        return 4 * self.B + 4 * self.B + 4 * self.B * self.slots * self.rw


def measure(arm: Arm, steps: int, warmup: int, barrier, allmax, clock_gpus=None):
    """Device-resident timing (CUDA events, max over ranks) + kernel counters, then end to end with pinned host buffers."""
    torch = arm.torch
    arm.run_steps(arm.step_dev, warmup)
    barrier()
    arm.ix.stats(reset=True)
    with ClockSampler(clock_gpus or [], enabled=bool(clock_gpus)) as clk:
        barrier()
        e0, e1 = arm.run_steps(arm.step_dev, steps)
        barrier()
    ms = e0.elapsed_time(e1)
    st = arm.ix.stats(reset=True)
    last = arm.d_out[(steps - 1) % arm.NS]
    flags_bad = int((last[1] != 0).sum().item())
    aligned = int((last[0] > 0).sum().item())
    ms_all = allmax(ms)
    arm.prepare_e2e()
    arm.run_steps(arm.step_e2e, max(1, warmup - 1))
    barrier()
    t0 = time.perf_counter()
    arm.run_steps(arm.step_e2e, steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_all = allmax(e2e_s)
    return {"ms": ms, "ms_ranks": ms_all, "stats": st, "aligned": aligned, "flags_bad": flags_bad, "e2e_s": e2e_s, "e2e_ranks": e2e_all,
            "e2e_aligned": int((arm.outs[(steps - 1) % arm.NS][0] > 0).sum()), "clocks": clk.summary() if clock_gpus else None}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=int(os.environ.get("BT_BENCH_READS", 8_000_000)))
    ap.add_argument("--streams", type=int, default=int(os.environ.get("BT_BENCH_STREAMS", 6)),
                    help="batches kept in flight (one bt_context_t + CUDA stream each), like the reference's -p worker threads")
    ap.add_argument("--policy", default=os.environ.get("BT_BENCH_POLICY", "n2k1"), choices=list(POLICIES),
                    help="headline policy; n2k1 also measures best and paired afterwards (other_policies) unless --no-others")
    ap.add_argument("--no-others", action="store_true", default=bool(os.environ.get("BT_BENCH_NO_OTHERS")))
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("BT_BENCH_CPU_SAMPLE", 1_000_000)))
    ap.add_argument("--cpu-step-seconds", type=float, default=float(os.environ.get("BT_BENCH_CPU_STEP_S", 6.0)))
    ap.add_argument("--mbp", type=int, default=int(os.environ.get("BT_BENCH_MBP", 3000)), help="size of the synthetic genome (Mbp)")
    ap.add_argument("--make-index", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--device", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.make_index:
        make_index(Path(args.make_index), args.mbp, args.device)
        return
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference" and rank != 0:
        return
    idx_info: dict = {}
    if args.policy == "v0" and not os.environ.get("BT_BENCH_INDEX"):
        os.environ["BT_BENCH_INDEX"] = str(REF_DIR / "fixtures" / "e_coli")          # BASELINE configs[1]: the reference's own index
    try:
        base, idx_name, idx_info = ensure_index(args.mbp, 0)
    except Exception as ex:
        base, idx_name = fallback_index()
        idx_info = {"fallback": f"hg19-sized index unavailable ({str(ex)[-300:]}); using {idx_name}"}
        print("bench.py: " + idx_info["fallback"], file=sys.stderr)
    if args.impl == "reference":
        reference_arm(args, base, idx_name, idx_info)
        return

    import torch
    import torch.distributed as dist
    import bowtie_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.environ.get("BOWTIE_B200_LIB"):
        bowtie_b200.build_library()
    t_load = time.time()
    ix = bowtie_b200.Index(str(base), need_mirror=True, device=local)
    t_load = time.time() - t_load
    genome = load_genome(base)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> list[float]:
        """Every rank's value (sorted); the job's time is the last."""
        if world == 1:
            return [float(x)]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return sorted(float(o.item()) for o in out)

    lib_comm = {"comm": None, "tried": False}

    def allreduce_counters(c: list[int]):
        """bt_counters_allreduce (the library's one collective) on a communicator made for it — the unique id travels through
        torch.distributed; falls back to dist.all_reduce if NCCL cannot be bound that way."""
        if world == 1:
            return c, "single rank"
        import ctypes as C
        if not lib_comm["tried"]:
            lib_comm["tried"] = True
            try:
                nccl = C.CDLL("libnccl.so.2")

                class UID(C.Structure):
                    _fields_ = [("internal", C.c_char * 128)]
                uid = UID()
                if rank == 0:
                    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
                t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).cuda()
                dist.broadcast(t, 0)
                C.memmove(C.byref(uid), bytes(t.cpu().numpy()), 128)
                comm = C.c_void_p()
                nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UID, C.c_int]
                assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
                lib_comm["comm"] = comm
            except Exception as ex:
                lib_comm["err"] = str(ex)[-200:]
        if lib_comm["comm"] is not None:
            L = bowtie_b200.load_library()
            arr = (C.c_uint64 * 5)(*c)
            L.bt_counters_allreduce.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
            if L.bt_counters_allreduce(lib_comm["comm"], arr, None) == 0:
                return [int(x) for x in arr], "bt_counters_allreduce (ncclAllReduce over 5 x u64)"
        t = torch.tensor(c, dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [int(x) for x in t.tolist()], "torch.distributed all_reduce (fallback: " + lib_comm.get("err", "library call failed") + ")"

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    def run_policy(name: str, B: int, NS: int, steps: int, warmup: int, headline: bool) -> dict:
        pd = POLICIES[name]
        arm = Arm(ix, name, genome, B, NS, rank, world, local)
        R = arm.R
        unit = "pairs/s" if R == 2 else "reads/s"
        m = measure(arm, steps, warmup, barrier, allmax, clock_gpus=list(range(world)) if (rank == 0 and headline) else None)
        ms_max, e2e_max = m["ms_ranks"][-1], m["e2e_ranks"][-1]
        value = B * steps * world / (ms_max / 1e3)
        st = m["stats"]
        alg = st.algorithmic_bytes / steps
        achieved = alg / (m["ms"] / steps / 1e3) / 1e9
        ctr = [m["aligned"], B - m["aligned"], 0, m["aligned"] * R if R == 1 else 0, m["aligned"] * R if R == 2 else 0]   # counters of the last step (hit.h:169-175)
        ctr, ctr_how = allreduce_counters(ctr)                                                          # the path's only collective
        # DRAM bytes per unit of the dominant kernel's launches, from the committed `ncu --set full` captures (profiles/): per step like `achieved`
        traffic, traffic_src = None, None
        try:
            tj = json.loads((ROOT / "profiles" / "r2_traffic_per_unit.json").read_text()).get(name)
            if tj:
                traffic, traffic_src = float(tj["dram_bytes_per_unit"]) * B, tj["source"]
        except Exception:
            pass
        res = {"metric": pd["metric"], "value": value, "unit": unit, "ms_per_step": ms_max / steps, "steps": steps, "warmup": warmup,
               "workload": workload_name(pd, idx_name), "units_per_step_per_gpu": B, "batches_in_flight": NS,
               "e2e": {"value": B * steps * world / e2e_max, "unit": unit, "h2d_bytes_per_step": arm.h2d, "d2h_bytes_per_step": arm.d2h},
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_step": alg,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                            "side_fetches_per_unit": st.side_fetches / (B * steps), "block_loads_per_unit": st.block_loads / (B * steps),
                            "algorithmic_bytes_per_unit": alg / B},
               "aligned_frac_last_step": m["aligned"] / B, "aligned_frac_last_e2e_step": m["e2e_aligned"] / B, "overflow_flags": m["flags_bad"],
               "counters_allreduced": [int(x) for x in ctr], "counters_collective": ctr_how,
               "rank_ms_per_step": {"min": m["ms_ranks"][0] / steps, "median": m["ms_ranks"][len(m["ms_ranks"]) // 2] / steps, "max": ms_max / steps},
               # per step: ctl_set x4, main search, collect, tail search, collect x2, ultra search, collect, overflow search
               # (best-first / paired: ctl_set x4, 4 arena tiers, 3 collects)
               # kernels of this repository inside the timed region, per step (one batch): DFS path = bt_ctl_set_all_kernel, bt_search_kernel (main),
               # bt_collect_kernel, bt_search_kernel (tail), bt_collect_kernel, bt_search_kernel (overflow); best-first / paired path =
               # bt_ctl_set_all_kernel, 4 x bt_best_kernel (arena tiers), 3 x bt_collect_kernel (profiles/r2_launches_*.csv)
               "gpu_launches": (6 if name in ("n2k1", "v0") else 8) * steps}
        if m["clocks"] is not None:
            res["clocks"] = m["clocks"]
        # latency of ONE synchronous batch through bt_align_batch (the INTEGRATION.md stub's call): host buffers in, host buffers out
        if headline and rank == 0:
            nl = min(B, 1_000_000)
            h = arm.host[0]
            sl = slice(0, nl * R * READ_LEN)
            ix.align(h[0][sl], h[1][sl], h[2][: nl * R + 1], h[3][: nl * R], arm.pol, slots=arm.slots, mm_cap=arm.mm_cap)   # warm (allocations)
            t0 = time.perf_counter()
            ix.align(h[0][sl], h[1][sl], h[2][: nl * R + 1], h[3][: nl * R], arm.pol, slots=arm.slots, mm_cap=arm.mm_cap)
            res["latency_ms_single_batch"] = {"ms": 1e3 * (time.perf_counter() - t0), "units": nl, "call": "bt_align_batch (synchronous, pageable host buffers)"}
        # the reference on the host cores + parity of the same reads (rank 0 of a single-GPU run only)
        if world == 1 and not os.environ.get("BT_BENCH_NO_CPU"):
            try:
                with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None) as tdn:
                    td = Path(tdn)
                    rr = ReferenceRunner(base, pd["flags"], R == 2, td, arm.host[0])
                    nmax = min(args.cpu_sample if headline else args.cpu_sample // 4, B)
                    rr.pick_threads(min(nmax, 200_000 if headline else 100_000))
                    n = int(min(nmax, max(50_000, rr.sweep[rr.threads] * args.cpu_step_seconds * (2 if headline else 1))))
                    (td / "st").mkdir()
                    fq = write_sample(td / "st", arm.host[0], n, R == 2)
                    r = rr.time_sample(fq, n, td / "ref.out")
                    res["cpu_baseline"] = {"value": n / r["search_s"], "unit": unit, "cores": rr.threads, "kind": "reference", "host_cores": rr.cores,
                                           "cpu_model": cpu_model(), "p_sweep": {str(k): round(v) for k, v in rr.sweep.items()},
                                           "sample": f"first {n} {'pairs' if R == 2 else 'reads'} of step 0, bowtie-align-s {' '.join(pd['flags'])} -p {rr.threads}; search-time basis = "
                                                     f"wall clock {r['wall_s']:.1f} s minus a one-read run {rr.overhead:.2f} s (process start + index load); 'Time searching' {r['time_searching_s']:.0f} s",
                                           "value_wall_clock": n / r["wall_s"]}
                    ref = parse_reference_output(td / "ref.out", [x.decode() if isinstance(x, bytes) else x for x in ix.refnames])
                    got = gpu_records(arm.outs[0][0], arm.outs[0][2], n, R == 2) if steps >= 1 else {}
                    # outs[0] holds the results of the last e2e step that used context 0; it is batch 0 iff that step index is even
                    last0 = max(k for k in range(steps) if k % arm.NS == 0)
                    if last0 & 1:
                        f, g, hh = ix.align(arm.host[0][0][: n * R * READ_LEN], arm.host[0][1][: n * R * READ_LEN], arm.host[0][2][: n * R + 1], arm.host[0][3][: n * R],
                                            arm.pol, slots=arm.slots, mm_cap=arm.mm_cap)
                        got = gpu_records(f, hh, n, R == 2)
                    res["parity_sample"] = {"units": n, **compare_parity(ref, got, r["gave_up"])}
                    if headline and name == "n2k1" and not os.environ.get("BT_BENCH_NO_CLI"):
                        try:
                            res["cli_e2e"] = cli_e2e(base, arm.host[0], min(B, 2_000_000), td, rr)
                        except Exception as ex:
                            res["cli_e2e"] = {"error": str(ex)[-300:]}
            except Exception as ex:  # the reference binary did not travel: report why instead of a number
                res["cpu_baseline"] = {"value": None, "unit": unit, "cores": host_cores(), "kind": "reference", "sample": f"unavailable: {str(ex)[-300:]}"}
        arm.close()
        return res

    B = args.reads_per_step
    NS = max(1, min(args.streams, args.steps))
    head = run_policy(args.policy, B, NS, args.steps, args.warmup, True)
    others = {}
    if args.policy == "n2k1" and not args.no_others:
        for name, b, ns in (("best", 2_000_000, 6), ("paired", 1_000_000, 6)):      # batch sizes from profiles/r2_kbench_call10.jsonl (the arenas are one pool per index)
            try:
                others[name] = run_policy(name, min(b, B), min(ns, args.steps), min(args.steps, 6), min(args.warmup, 3), False)
            except Exception as ex:
                others[name] = {"error": str(ex)[-300:]}
    # strong scaling (config 4 is a fixed read set): the same 8 batches of reads split over the ranks
    strong = None
    if args.policy == "n2k1" and not os.environ.get("BT_BENCH_NO_STRONG"):
        total_batches = 8
        per_rank = max(1, total_batches // world)
        arm = Arm(ix, "n2k1", genome, B, min(NS, per_rank), rank, world, local)
        arm.run_steps(arm.step_dev, 1)
        barrier()
        e0, e1 = arm.run_steps(arm.step_dev, per_rank)
        barrier()
        ms = allmax(e0.elapsed_time(e1))
        strong = {"total_reads": B * per_rank * world, "reads_per_s": B * per_rank * world / (ms[-1] / 1e3), "batches_per_rank": per_rank, "scaling": "strong"}
        arm.close()

    if rank == 0:
        line = {
            "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": head["workload"], "reads_per_step_per_gpu": B, "index_len_bp": ix.len, "index_device_bytes": ix.device_bytes,
                       "index": idx_info, "index_load_s": round(t_load, 1),
                       "parallelism": f"reads sharded over {world} GPU(s), full index per GPU; {NS} batches in flight per GPU (contexts/streams)",
                       "l2": f"two alternating read batches of {2 * B * READ_LEN / 1e6:.0f} MB each (> 126 MB L2); the index ({ix.device_bytes / 1e6:.0f} MB on the device) "
                             + ("exceeds L2 too" if ix.device_bytes > 126e6 else "is L2-resident (the reference's own e_coli index)"),
                       "aligned_frac_last_step": head["aligned_frac_last_step"], "aligned_frac_last_e2e_step": head["aligned_frac_last_e2e_step"],
                       "overflow_flags": head["overflow_flags"], "counters_allreduced": head["counters_allreduced"],
                       "rank_ms_per_step": head["rank_ms_per_step"]},
            "clocks": head.get("clocks"), "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "cpu_baseline": head.get("cpu_baseline"),
        }
        for k in ("parity_sample", "latency_ms_single_batch", "cli_e2e"):
            if k in head:
                line[k] = head[k]
        if others:
            line["other_policies"] = others
        if strong:
            line["strong_scaling"] = strong
        print(json.dumps(line))
        bad = [("headline", head)] + list(others.items())
        for nm, r in bad:
            ps = r.get("parity_sample") if isinstance(r, dict) else None
            if ps and ps.get("mismatching"):
                print(f"bench.py: PARITY FAILURE in {nm}: {ps}", file=sys.stderr)
                if world > 1:
                    dist.destroy_process_group()
                sys.exit(3)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
