#!/usr/bin/env python
"""bench.py — aligned reads/s of the B200 FM-index backward-search path (100 bp, -n 2).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
    python bench.py --impl reference [--gpus N] ...               the reference's CPU path (oracle/_ref)

A "step" is one pass of the hot path (bt_align_batch*, all phases of `-n 2 -k 1`) over one batch of
synthetic 100-bp reads.  `value` = whole-job reads/s with the batch already resident in HBM; `e2e` = the
same through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region).  `roofline` is the
search kernel's algorithmic bytes (64 B per side fetch + 4 B per offs[] read + 8 B per ftab
jump, counted by the kernel itself) over its device time.  Under torchrun every rank owns one GPU, holds the
whole index and a disjoint shard of reads (weak scaling); the only collective is the all-reduce of the five
hit counters.
"""
from __future__ import annotations

import argparse
import json
import os
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


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------

def pick_index() -> tuple[Path, str]:
    """Largest prebuilt synthetic index under oracle/_ref/cache (built once with the reference's
    bowtie-build; see tests/synth.py), else the reference's shipped e_coli index."""
    env = os.environ.get("BT_BENCH_INDEX")
    if env:
        return Path(env), Path(env).name
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
    raise SystemExit("bench.py: no index available (expected oracle/_ref/cache/bench_* or oracle/_ref/fixtures/e_coli)")


def load_genome(base: Path) -> np.ndarray:
    """Base codes 0..3 of the joined reference from X.4.ebwt (2-bit packed, base i in bits 2*(i&3) of byte i>>2;
    reference.h:58-330).  Reads are sampled from this text, so no FASTA has to travel with the index."""
    raw = np.frombuffer(Path(str(base) + ".4.ebwt").read_bytes(), np.uint8)
    out = np.empty(len(raw) * 4, np.uint8)
    for k in range(4):
        out[k::4] = (raw >> (2 * k)) & 3
    return out


def make_reads(genome: np.ndarray, n: int, seed: int):
    """n x 100 bp: uniform positions, both strands, 1 % substitutions, 0.1 % random reads, Phred33 quals from
    {40,40,40,35,30,20,10}.  Returns (codes[n*100], quals[n*100], offs[n+1], seeds[n], names)."""
    rng = np.random.default_rng(seed)
    L = READ_LEN
    pos = rng.integers(0, len(genome) - L, n)
    codes = genome[pos[:, None] + np.arange(L)[None, :]]
    mut = rng.random((n, L)) < 0.01
    codes[mut] = (codes[mut] + rng.integers(1, 4, int(mut.sum()))) & 3
    rnd = rng.random(n) < 0.001
    codes[rnd] = rng.integers(0, 4, (int(rnd.sum()), L))
    rc = rng.random(n) < 0.5
    flipped = codes[rc, ::-1]
    codes[rc] = np.where(flipped < 4, 3 - flipped, 4)
    codes = np.ascontiguousarray(codes, np.uint8)
    quals = (rng.choice(np.array([40, 40, 40, 35, 30, 20, 10], np.uint8), (n, L)) + 33).astype(np.uint8)
    # names "r%09d" (fixed width so genRandSeed vectorises); Read::seed per pat.cpp:21-57 with global seed 0
    ids = np.arange(n, dtype=np.uint32)
    name = np.zeros((n, 10), np.uint8)
    name[:, 0] = ord("r")
    for d in range(9):
        name[:, 9 - d] = (ids // 10 ** d) % 10 + 48
    seeds = np.full(n, ((0 + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83) & 0xFFFFFFFF, np.uint32)
    i = np.arange(L)
    seeds ^= np.bitwise_xor.reduce(codes.astype(np.uint32) << ((i & 15) << 1).astype(np.uint32), axis=1)
    seeds ^= np.bitwise_xor.reduce(quals.astype(np.uint32) << ((i & 3) << 3).astype(np.uint32), axis=1)
    j = np.arange(10)
    seeds ^= np.bitwise_xor.reduce(name.astype(np.uint32) << ((j & 3) << 3).astype(np.uint32), axis=1)
    offs = (np.arange(n + 1, dtype=np.uint64) * L)
    return codes.reshape(-1), quals.reshape(-1), offs, seeds.astype(np.uint32), name


def make_pairs(genome: np.ndarray, n: int, seed: int):
    """n pairs of 2 x 100 bp in --fr geometry: fragment length N(200, 20) clipped to [101, 250], mate 1 the
    fragment's first 100 bases, mate 2 the reverse complement of its last 100, the whole fragment flipped half of the time,
    1 % substitutions.  Reads are interleaved (mate 1, mate 2, mate 1, ...); names "r%09d/1", "r%09d/2"."""
    rng = np.random.default_rng(seed)
    L = READ_LEN
    F = np.clip(rng.normal(200.0, 20.0, n).round().astype(np.int64), L + 1, 250)
    pos = rng.integers(0, len(genome) - 251, n)
    left = genome[pos[:, None] + np.arange(L)[None, :]]
    right = genome[(pos + F - L)[:, None] + np.arange(L)[None, :]]
    rcomp = lambda a: np.where(a[:, ::-1] < 4, 3 - a[:, ::-1], 4)
    flip = rng.random(n) < 0.5
    m1 = np.where(flip[:, None], rcomp(right), left)
    m2 = np.where(flip[:, None], left, rcomp(right))
    codes = np.empty((2 * n, L), np.uint8)
    codes[0::2] = m1; codes[1::2] = m2
    mut = rng.random((2 * n, L)) < 0.01
    codes[mut] = (codes[mut] + rng.integers(1, 4, int(mut.sum()))) & 3
    quals = (rng.choice(np.array([40, 40, 40, 35, 30, 20, 10], np.uint8), (2 * n, L)) + 33).astype(np.uint8)
    ids = np.repeat(np.arange(n, dtype=np.uint32), 2)
    name = np.zeros((2 * n, 12), np.uint8)
    name[:, 0] = ord("r")
    for d in range(9):
        name[:, 9 - d] = (ids // 10 ** d) % 10 + 48
    name[:, 10] = ord("/"); name[0::2, 11] = ord("1"); name[1::2, 11] = ord("2")
    seeds = np.full(2 * n, ((0 + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83) & 0xFFFFFFFF, np.uint32)
    i = np.arange(L)
    seeds ^= np.bitwise_xor.reduce(codes.astype(np.uint32) << ((i & 15) << 1).astype(np.uint32), axis=1)
    seeds ^= np.bitwise_xor.reduce(quals.astype(np.uint32) << ((i & 3) << 3).astype(np.uint32), axis=1)
    j = np.arange(12)
    seeds ^= np.bitwise_xor.reduce(name.astype(np.uint32) << ((j & 3) << 3).astype(np.uint32), axis=1)
    offs = (np.arange(2 * n + 1, dtype=np.uint64) * L)
    return codes.reshape(-1), quals.reshape(-1), offs, seeds.astype(np.uint32), name


def write_fastq_pairs(p1: Path, p2: Path, codes, quals, name, npairs: int) -> None:
    L = READ_LEN
    lut = np.frombuffer(b"ACGTN", np.uint8)
    seq = lut[codes[: 2 * npairs * L]].reshape(2 * npairs, L)
    q = quals[: 2 * npairs * L].reshape(2 * npairs, L)
    W = name.shape[1]
    rec = np.empty((2 * npairs, 1 + W + 1 + L + 3 + L + 1), np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1:1 + W] = name[: 2 * npairs]; rec[:, 1 + W] = 10
    rec[:, 2 + W:2 + W + L] = seq; rec[:, 2 + W + L] = 10; rec[:, 3 + W + L] = ord("+"); rec[:, 4 + W + L] = 10
    rec[:, 5 + W + L:5 + W + 2 * L] = q; rec[:, 5 + W + 2 * L] = 10
    p1.write_bytes(rec[0::2].tobytes()); p2.write_bytes(rec[1::2].tobytes())


def write_fastq(path: Path, codes, quals, name, n: int) -> None:
    L = READ_LEN
    lut = np.frombuffer(b"ACGTN", np.uint8)
    seq = lut[codes[: n * L]].reshape(n, L)
    q = quals[: n * L].reshape(n, L)
    rec = np.empty((n, 1 + 10 + 1 + L + 3 + L + 1), np.uint8)
    rec[:, 0] = ord("@"); rec[:, 1:11] = name[:n]; rec[:, 11] = 10
    rec[:, 12:12 + L] = seq; rec[:, 12 + L] = 10; rec[:, 13 + L] = ord("+"); rec[:, 14 + L] = 10
    rec[:, 15 + L:15 + 2 * L] = q; rec[:, 15 + 2 * L] = 10
    path.write_bytes(rec.tobytes())


# ------------------------------------------------------------------------------------------------
# clocks (recipe of /opt/skills/guides/B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu: int) -> None:
        self.gpu, self.samples, self.stop = gpu, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self) -> None:
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].startswith("Active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference binary on the host cores
# ------------------------------------------------------------------------------------------------

def run_reference_sample(base: Path, fq, n: int, threads: int, flags=("-n", "2", "-k", "1")) -> tuple[float, float]:
    """bowtie-align-s -n 2 -k 1 -t -p <threads>; returns (search seconds from -t, wall seconds)."""
    exe = REF_DIR / "bowtie-align-s"
    if not exe.exists():
        raise RuntimeError("oracle/_ref/bowtie-align-s missing (built by oracle/Makefile from /root/reference)")
    t0 = time.time()
    inputs = ["-1", str(fq[0]), "-2", str(fq[1])] if isinstance(fq, (tuple, list)) else [str(fq)]      # a pair of mate files, or one file
    p = subprocess.run([str(exe), *flags, "-t", "-p", str(threads), "-x", str(base), *inputs, "/dev/null"],
                       capture_output=True, text=True)
    wall = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError("reference run failed: " + p.stderr[-500:])
    search = None
    for line in (p.stdout + p.stderr).splitlines():
        if line.startswith("Time searching:"):
            h, m, s = line.split(":", 1)[1].strip().split(":")
            search = int(h) * 3600 + int(m) * 60 + int(s)
    return (float(search) if search else wall), wall


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=int(os.environ.get("BT_BENCH_READS", 4_000_000)))
    ap.add_argument("--streams", type=int, default=int(os.environ.get("BT_BENCH_STREAMS", 8)),
                    help="batches kept in flight (one bt_context_t + CUDA stream each), like the reference's -p worker threads")
    ap.add_argument("--policy", default=os.environ.get("BT_BENCH_POLICY", "n2k1"), choices=["n2k1", "best", "paired", "v0"],
                    help="n2k1: the headline workload (-n 2 -k 1, SURVEY config 4); best: -n 2 --best (config 3, best-first path); "
                         "paired: -n 3 on 2x100 bp pairs (config 5; value counts PAIRS per second); v0: -v 0 exact on the shipped e_coli index (config 2)")
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("BT_BENCH_CPU_SAMPLE", 1_000_000)))
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.policy == "v0" and not os.environ.get("BT_BENCH_INDEX"):
        os.environ["BT_BENCH_INDEX"] = str(REF_DIR / "fixtures" / "e_coli")          # BASELINE configs[1]: the reference's own index
    base, idx_name = pick_index()
    ref_flags = {"n2k1": ["-n", "2", "-k", "1"], "best": ["-n", "2", "--best"], "paired": ["-n", "3"], "v0": ["-v", "0"]}[args.policy]
    R = 2 if args.policy == "paired" else 1                     # reads per work unit
    unit = "pairs/s" if R == 2 else "reads/s"
    metric = "aligned read pairs/sec (2x100 bp, -n 3 paired-end)" if R == 2 else "aligned reads/sec (100 bp, -v 0)" if args.policy == "v0" else "aligned reads/sec (100 bp, -n 2)"
    gen = make_pairs if R == 2 else make_reads
    cfg_workload = f"{' '.join(ref_flags)}, {'2x' if R == 2 else ''}{READ_LEN} bp synthetic {'pairs (--fr, fragments N(200,20))' if R == 2 else 'reads'} (1% subs, both strands), index {idx_name}"

    def write_sample(td: Path, h, n: int):
        if R == 2:
            write_fastq_pairs(td / "s_1.fq", td / "s_2.fq", h[0], h[1], h[4], n)
            return (td / "s_1.fq", td / "s_2.fq")
        write_fastq(td / "s.fq", h[0], h[1], h[4], n)
        return td / "s.fq"
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        genome = load_genome(base)
        n = min(args.cpu_sample, args.reads_per_step)
        h0 = gen(genome, n, seed=12345)
        with tempfile.TemporaryDirectory() as td:
            fq = write_sample(Path(td), h0, n)
            times = []
            for it in range(args.warmup + args.steps):
                search, wall = run_reference_sample(base, fq, n, cores, ref_flags)
                if it >= args.warmup:
                    times.append(wall)
        tot = sum(times)
        val = n * args.steps / tot
        line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": cfg_workload, "reads_per_step": n, "parallelism": f"{cores} host threads (bowtie -p)"},
                "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "reference",
                                 "sample": f"{n} reads per step, wall clock of bowtie-align-s {' '.join(ref_flags)} -p {cores} incl. index load"},
                "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
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
    ix = bowtie_b200.Index(str(base), need_mirror=True, device=local)
    pol = bowtie_b200.Policy(mode=1, mms=3 if R == 2 else 2, khits=1, best=(args.policy == "best"), paired=(R == 2))
    if args.policy == "v0":
        pol = bowtie_b200.Policy(mode=0, mms=0, khits=1)
    if args.policy in ("best", "paired"):
        args.streams = min(args.streams, 3)      # every context of the best-first path owns ~12 GB of arenas
    B, L, slots, mm_cap = args.reads_per_step, READ_LEN, R, 7
    rw = bowtie_b200.BT_HIT_HDR_WORDS + mm_cap
    genome = load_genome(base)
    # two distinct batches per rank, alternated, each larger than L2 (4M reads x 200 B = 800 MB)
    host = [gen(genome, B, seed=12345 + 1000 * rank + k) for k in range(2)]
    dev = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (h[0], h[1], h[2].view(np.int64), h[3].view(np.int32))) for h in host]
    NS = max(1, min(args.streams, args.steps))
    ctxs = [bowtie_b200.Context(ix) for _ in range(NS)]
    streams = [torch.cuda.Stream() for _ in range(NS)]
    d_out = [(torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda"),
              torch.zeros(B * slots * rw, dtype=torch.int32, device="cuda")) for _ in range(NS)]
    main = torch.cuda.current_stream()

    def step_dev(k: int) -> None:
        s, q, o, sd = dev[k & 1]
        f, g, h = d_out[k % NS]
        ctxs[k % NS].align_device(s.data_ptr(), q.data_ptr(), o.data_ptr(), sd.data_ptr(), B * R, L, pol, f.data_ptr(), g.data_ptr(),
                                  h.data_ptr(), slots, mm_cap, streams[k % NS].cuda_stream)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(fn, nsteps: int):
        """K steps round-robin over NS streams, bracketed by events on the main stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        for k in range(nsteps):
            fn(k)
        for cx, st in zip(ctxs, streams):
            cx.join(st.cuda_stream)            # the batch's heavy / overflow passes (and D2H) run on the context's side stream
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        return e0, e1

    run_steps(step_dev, args.warmup)
    barrier()
    ix.stats(reset=True)
    with ClockSampler(local) as clk:
        barrier()
        e0, e1 = run_steps(step_dev, args.steps)
        barrier()
    ms = e0.elapsed_time(e1)
    st = ix.stats(reset=True)
    d_found, d_flags = d_out[(args.steps - 1) % NS][0], d_out[(args.steps - 1) % NS][1]
    flags_bad = int((d_flags != 0).sum().item())
    aligned = int((d_found > 0).sum().item())
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    ctr = torch.tensor([aligned, B - aligned, 0, aligned, 0], dtype=torch.int64, device="cuda")   # counters of the last step (hit.h:169-175)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ctr, op=dist.ReduceOp.SUM)                                                  # the path's only collective
    ms_max = float(t.item())
    value = B * args.steps * world / (ms_max / 1e3)

    # end to end through the host-buffer entry point: pinned host inputs, H2D + kernels + D2H every step
    pin = []
    for h in host:
        ts = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (h[0], h[1], h[2].view(np.int64), h[3].view(np.int32))]
        pin.append(ts)
    outs = []
    for _ in range(NS):
        o_found = torch.zeros(B, dtype=torch.int32).pin_memory()
        o_flags = torch.zeros(B, dtype=torch.int32).pin_memory()
        o_hits = torch.zeros(B * slots * rw, dtype=torch.int32).pin_memory()
        outs.append((o_found.numpy().view(np.uint32), o_flags.numpy().view(np.uint32), o_hits.numpy().view(np.uint32).reshape(B, slots, rw)))

    def step_e2e(k: int) -> None:
        s, q, o, sd = pin[k & 1]
        ctxs[k % NS].align_async(s.numpy(), q.numpy(), o.numpy().view(np.uint64), sd.numpy().view(np.uint32), pol, outs[k % NS],
                                 slots, mm_cap, streams[k % NS].cuda_stream)

    run_steps(step_e2e, max(1, args.warmup - 1))
    barrier()
    t0 = time.perf_counter()
    run_steps(step_e2e, args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_aligned = int((outs[(args.steps - 1) % NS][0] > 0).sum())
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = B * args.steps * world / float(t.item())
    h2d = 2 * B * R * L + 8 * (B * R + 1) + 4 * B * R
    d2h = 4 * B + 4 * B + 4 * B * slots * rw

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = st.algorithmic_bytes / args.steps
    achieved = alg_bytes / (ms / args.steps / 1e3) / 1e9
    traffic = None
    tp = ROOT / "profiles" / "traffic_per_launch.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            pass
    cpu = None
    if world == 1 and not os.environ.get("BT_BENCH_NO_CPU"):
        try:
            n = min(args.cpu_sample, B)
            with tempfile.TemporaryDirectory() as td:
                fq = write_sample(Path(td), host[0], n)
                search, wall = run_reference_sample(base, fq, n, cores, ref_flags)
            cpu = {"value": n / wall, "unit": unit, "cores": cores, "kind": "reference",
                   "sample": f"first {n} reads of step 0, bowtie-align-s {' '.join(ref_flags)} -p {cores}, wall clock {wall:.1f}s incl. index load ('Time searching' {search:.0f}s)"}
        except Exception as ex:  # the reference binary did not travel: report why instead of a number
            cpu = {"value": None, "unit": unit, "cores": cores, "kind": "reference", "sample": f"unavailable: {ex}"}
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": cfg_workload, "reads_per_step_per_gpu": B, "index_len_bp": ix.len, "index_device_bytes": ix.device_bytes,
                   "parallelism": f"reads sharded over {world} GPU(s), full index per GPU; {NS} batches in flight per GPU (contexts/streams)",
                   "l2": f"two alternating read batches of {2 * B * L / 1e6:.0f} MB each (> 126 MB L2); the index ({ix.device_bytes / 1e6:.0f} MB on the device) "
                         + ("exceeds L2 too" if ix.device_bytes > 126e6 else "is L2-resident (the reference's own e_coli index)"),
                   "aligned_frac_last_step": aligned / B, "aligned_frac_last_e2e_step": e2e_aligned / B, "overflow_flags": flags_bad,
                   "counters_allreduced": [int(x) for x in ctr.tolist()]},
        "clocks": clk.summary(),
        "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        # per step: 3 ctl_set, main search, 3 collect, heavy search, overflow search (best-first / paired: 4 ctl_set, 4 arena tiers, 3 collect)
        "gpu_launches": (9 if args.policy in ("n2k1", "v0") else 11) * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                     "side_fetches_per_read": st.side_fetches / (B * args.steps), "block_loads_per_read": st.block_loads / (B * args.steps),
                     "algorithmic_bytes_per_read": alg_bytes / B},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
