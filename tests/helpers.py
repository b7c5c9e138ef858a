"""Shared test infrastructure: read parsing, oracle (ctypes) binding, Bowtie default-format
rendering, and a runner for the compiled reference binary (oracle/_ref/bowtie-align-s).

Nothing here is product code; the product lives in bowtie_b200/.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"
FIXTURES = REF_DIR / "fixtures"
REF_ALIGN = REF_DIR / "bowtie-align-s"
REF_BUILD = REF_DIR / "bowtie-build-s"
GOLDEN = ROOT / "tests" / "golden"

ASC2DNA = np.full(256, 4, dtype=np.uint8)
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    ASC2DNA[ord(_ch)] = _v
    ASC2DNA[ord(_ch.lower())] = _v


def ensure_oracle_built() -> None:
    """Build oracle/libbtoracle.so (and oracle/_ref when /root/reference exists)."""
    so = ORACLE_DIR / "libbtoracle.so"
    src = ORACLE_DIR / "bt_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "port"], check=True, capture_output=True)
    if not REF_ALIGN.exists() and Path("/root/reference/ebwt_search.cpp").exists():
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-j8", "ref"], check=True, capture_output=True)
    # the fixtures are copies of the reference's shipped files; a copy that was clobbered (e.g. by handing it to an aligner as
    # an output path) is restored where the originals are available, and reported where they are not
    for f in FIXTURES.glob("*") if FIXTURES.exists() else []:
        if f.is_file() and f.stat().st_size == 0:
            for d in ("reads", "indexes", "genomes"):
                src = Path("/root/reference") / d / f.name
                if src.exists():
                    import shutil
                    shutil.copyfile(src, f)
                    break
            else:
                raise RuntimeError(f"fixture {f} is empty and /root/reference is not available to restore it")


def have_reference() -> bool:
    return REF_ALIGN.exists() and (FIXTURES / "e_coli.1.ebwt").exists()


# ----------------------------------------------------------------------------------------
# reads
# ----------------------------------------------------------------------------------------

@dataclass
class ReadBatch:
    names: list[bytes]
    seqs: list[bytes]          # ASCII, upper-cased as the reference prints them (ACGTN)
    quals: list[bytes]         # phred+33 chars
    seq_codes: np.ndarray = field(default=None)    # concatenated 0..4
    qual_cat: np.ndarray = field(default=None)     # concatenated phred+33
    offs: np.ndarray = field(default=None)         # uint64 [n+1]
    seeds: np.ndarray = field(default=None)        # uint32 [n]

    def __len__(self) -> int:
        return len(self.names)


def gen_rand_seed(codes: np.ndarray, qual: bytes, name: bytes, global_seed: int = 0) -> int:
    """pat.cpp:21-57 genRandSeed."""
    rseed = ((global_seed + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83) & 0xFFFFFFFF
    for i, p in enumerate(codes.tolist()):
        rseed ^= (p << ((i & 15) << 1)) & 0xFFFFFFFF
    for i, p in enumerate(qual):
        rseed ^= (p << ((i & 3) << 3)) & 0xFFFFFFFF
    for i, p in enumerate(name):
        rseed ^= (p << ((i & 3) << 3)) & 0xFFFFFFFF
    return rseed & 0xFFFFFFFF


def finalize_batch(names, seqs, quals, global_seed: int = 0) -> ReadBatch:
    codes = [ASC2DNA[np.frombuffer(s, dtype=np.uint8)] if len(s) else np.zeros(0, np.uint8) for s in seqs]
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
    seq_codes = np.concatenate(codes) if codes else np.zeros(0, np.uint8)
    qual_cat = np.frombuffer(b"".join(quals), dtype=np.uint8).copy() if quals else np.zeros(0, np.uint8)
    seeds = np.array([gen_rand_seed(c, q, n, global_seed) for c, q, n in zip(codes, quals, names)], dtype=np.uint32)
    # the reference prints sequences from the decoded codes: anything not ACGT becomes N
    norm = [bytes(b"ACGTN"[v] for v in c.tolist()) for c in codes]
    return ReadBatch(list(names), norm, list(quals), seq_codes, qual_cat, offs, seeds)


def parse_fastq(path: str | os.PathLike, global_seed: int = 0) -> ReadBatch:
    """FastqPatternSource::parse (pat.cpp:862-975), phred33, no trimming."""
    names, seqs, quals = [], [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    rdid = 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        assert lines[i][:1] == b"@", lines[i]
        name = lines[i][1:].rstrip(b"\r")
        seq = bytes(ch for ch in lines[i + 1].replace(b".", b"N") if chr(ch).isalpha())
        qual = lines[i + 3].rstrip(b"\r")
        if not name:
            name = str(rdid).encode()
        names.append(name)
        seqs.append(seq)
        quals.append(qual)
        rdid += 1
        i += 4
    return finalize_batch(names, seqs, quals, global_seed)


# ----------------------------------------------------------------------------------------
# oracle binding
# ----------------------------------------------------------------------------------------

class _Policy(C.Structure):
    _fields_ = [("mode", C.c_int), ("mms", C.c_int), ("seedLen", C.c_int), ("qualThresh", C.c_int),
                ("maxBts", C.c_int), ("khits", C.c_uint32), ("mhits", C.c_uint32), ("allHits", C.c_int),
                ("nofw", C.c_int), ("norc", C.c_int), ("maqRound", C.c_int)]


class _Hit(C.Structure):
    _fields_ = [("read", C.c_uint32), ("tidx", C.c_uint32), ("toff", C.c_uint32), ("oms", C.c_uint32),
                ("cost", C.c_uint16), ("fw", C.c_uint8), ("stratum", C.c_uint8), ("nmm", C.c_uint32),
                ("mm_off", C.c_uint32)]


class _Mm(C.Structure):
    _fields_ = [("pos", C.c_uint16), ("refc", C.c_uint8), ("pad", C.c_uint8)]


class _Stats(C.Structure):
    _fields_ = [("lfex", C.c_uint64), ("lf", C.c_uint64), ("chase", C.c_uint64), ("ftab", C.c_uint64),
                ("offs", C.c_uint64), ("backtracks", C.c_uint64)]


class _Result(C.Structure):
    _fields_ = [("hits", C.POINTER(_Hit)), ("nhits", C.c_size_t), ("cap_hits", C.c_size_t),
                ("mms", C.POINTER(_Mm)), ("nmms", C.c_size_t), ("cap_mms", C.c_size_t),
                ("nhits_per_read", C.POINTER(C.c_uint32)), ("maxed", C.POINTER(C.c_uint8)),
                ("counters", C.c_uint64 * 5), ("stats", _Stats)]


HIT_DTYPE = np.dtype([("read", "<u4"), ("tidx", "<u4"), ("toff", "<u4"), ("oms", "<u4"), ("cost", "<u2"),
                      ("fw", "u1"), ("stratum", "u1"), ("nmm", "<u4"), ("mm_off", "<u4")])
MM_DTYPE = np.dtype([("pos", "<u2"), ("refc", "u1"), ("pad", "u1")])


@dataclass
class Policy:
    mode: int = 1            # 0: -v, 1: -n
    mms: int = 2
    seed_len: int = 28
    qual_thresh: int = 70
    max_bts: int = 125
    khits: int = 1
    mhits: int = 0xFFFFFFFF
    all_hits: bool = False
    nofw: bool = False
    norc: bool = False
    maq_round: bool = True

    def to_c(self) -> _Policy:
        return _Policy(self.mode, self.mms, self.seed_len, self.qual_thresh, self.max_bts, self.khits,
                       self.mhits, int(self.all_hits), int(self.nofw), int(self.norc), int(self.maq_round))

    def ref_args(self) -> list[str]:
        a = ["-v" if self.mode == 0 else "-n", str(self.mms)]
        if self.mode == 1:
            a += ["-l", str(self.seed_len), "-e", str(self.qual_thresh), "--maxbts", str(self.max_bts)]
            if not self.maq_round:
                a.append("--nomaqround")
        if self.all_hits:
            a.append("-a")
        else:
            a += ["-k", str(self.khits)]
        if self.mhits != 0xFFFFFFFF:
            a += ["-m", str(self.mhits)]
        if self.nofw:
            a.append("--nofw")
        if self.norc:
            a.append("--norc")
        return a


@dataclass
class AlignResult:
    hits: np.ndarray            # HIT_DTYPE, in report order (read-major)
    mms: np.ndarray             # MM_DTYPE
    nhits_per_read: np.ndarray
    maxed: np.ndarray
    counters: np.ndarray        # aligned, unaligned, maxed, reported, reportedPaired
    stats: dict | None = None


class Oracle:
    """ctypes binding of oracle/libbtoracle.so (the CPU restatement)."""

    def __init__(self) -> None:
        ensure_oracle_built()
        L = C.CDLL(str(ORACLE_DIR / "libbtoracle.so"))
        L.bto_index_load.restype = C.c_void_p
        L.bto_index_load.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.bto_index_free.argtypes = [C.c_void_p]
        L.bto_result_new.restype = C.POINTER(_Result)
        L.bto_result_new.argtypes = [C.c_size_t]
        L.bto_result_free.argtypes = [C.POINTER(_Result)]
        L.bto_align.restype = C.c_int
        L.bto_align.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Policy), C.c_size_t, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.POINTER(_Result)]
        L.bto_map_lf.restype = C.c_uint32
        L.bto_map_lf.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.bto_map_lf1.restype = C.c_uint32
        L.bto_map_lf1.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.bto_map_lf_ex.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.bto_row_l.restype = C.c_int
        L.bto_row_l.argtypes = [C.c_void_p, C.c_uint32]
        L.bto_chase.restype = C.c_uint32
        L.bto_chase.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.bto_ftab_hi.restype = C.c_uint32
        L.bto_ftab_hi.argtypes = [C.c_void_p, C.c_uint32]
        L.bto_ftab_lo.restype = C.c_uint32
        L.bto_ftab_lo.argtypes = [C.c_void_p, C.c_uint32]
        L.bto_gen_rand_seed.restype = C.c_uint32
        L.bto_gen_rand_seed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32]
        self.L = L
        self._idx: dict[tuple[str, int], int] = {}

    def index(self, base: str | os.PathLike, mirror: bool) -> int:
        key = (str(base), int(mirror))
        if key not in self._idx:
            err = C.create_string_buffer(512)
            h = self.L.bto_index_load(str(base).encode(), int(mirror), err, 512)
            if not h:
                raise RuntimeError(err.value.decode())
            self._idx[key] = h
        return self._idx[key]

    def align(self, base, batch: ReadBatch, pol: Policy) -> AlignResult:
        fw = self.index(base, False)
        need_mirror = pol.mode == 1 or pol.mms > 0
        bw = self.index(base, True) if need_mirror else None
        res = self.L.bto_result_new(len(batch))
        try:
            cpol = pol.to_c()
            seq = np.ascontiguousarray(batch.seq_codes)
            qual = np.ascontiguousarray(batch.qual_cat)
            rc = self.L.bto_align(fw, bw, C.byref(cpol), len(batch), seq.ctypes.data, qual.ctypes.data,
                                  batch.offs.ctypes.data, batch.seeds.ctypes.data, res)
            if rc != 0:
                raise RuntimeError(f"bto_align rc={rc}")
            r = res.contents
            n = len(batch)
            hits = np.ctypeslib.as_array(C.cast(r.hits, C.POINTER(C.c_uint8)), (r.nhits * HIT_DTYPE.itemsize,)).view(HIT_DTYPE).copy() \
                if r.nhits else np.zeros(0, HIT_DTYPE)
            mms = np.ctypeslib.as_array(C.cast(r.mms, C.POINTER(C.c_uint8)), (r.nmms * MM_DTYPE.itemsize,)).view(MM_DTYPE).copy() \
                if r.nmms else np.zeros(0, MM_DTYPE)
            nh = np.ctypeslib.as_array(r.nhits_per_read, (max(n, 1),))[:n].copy()
            mx = np.ctypeslib.as_array(r.maxed, (max(n, 1),))[:n].copy()
            ctr = np.array(list(r.counters), dtype=np.uint64)
            st = {k: int(getattr(r.stats, k)) for k, _ in _Stats._fields_}
            return AlignResult(hits, mms, nh, mx, ctr, st)
        finally:
            self.L.bto_result_free(res)


# ----------------------------------------------------------------------------------------
# rendering: Bowtie default output (hit.cpp:73-301, VerboseHitSink::append)
# ----------------------------------------------------------------------------------------

_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def refname_upto_ws(name: bytes) -> bytes:
    return name.split()[0] if name.split() else name


def load_refnames(base) -> list[bytes]:
    """Reference names from X.1.ebwt (ebwt.h:3258-3272)."""
    with open(f"{base}.1.ebwt", "rb") as f:
        d = f.read()
    w = np.frombuffer(d[:28], dtype="<u4")
    length, ftab_chars = int(w[1]), int(np.frombuffer(d[20:24], dtype="<i4")[0])
    pos = 28
    npat = int(np.frombuffer(d[pos:pos + 4], "<u4")[0]); pos += 4 + 4 * npat
    nfrag = int(np.frombuffer(d[pos:pos + 4], "<u4")[0]); pos += 4 + 12 * nfrag
    bwt_sz = length // 4 + 1
    pos += ((bwt_sz + 111) // 112) * 128
    pos += 4 + 20 + 4 * ((1 << (2 * ftab_chars)) + 1) + 4 * 2 * ftab_chars
    raw = d[pos:].split(b"\0")[0]
    names = raw.split(b"\n")
    if names and names[-1] == b"":
        names = names[:-1]
    return names


def render_default(batch: ReadBatch, res: AlignResult, refnames: list[bytes]) -> bytes:
    out = []
    for h in res.hits:
        i = int(h["read"])
        fw = bool(h["fw"])
        seq = batch.seqs[i]
        qual = batch.quals[i]
        if not fw:
            seq = seq.translate(_COMP)[::-1]
            qual = qual[::-1]
        t = int(h["tidx"])
        rn = refname_upto_ws(refnames[t]) if t < len(refnames) else str(t).encode()
        mm = res.mms[int(h["mm_off"]): int(h["mm_off"]) + int(h["nmm"])]
        ents = sorted((int(m["pos"]), int(m["refc"])) for m in mm)
        L = len(seq)
        parts = []
        for pos, refc in ents:
            qch = seq[pos] if fw else seq[L - pos - 1]
            parts.append(b"%d:%c>%c" % (pos, b"ACGT"[refc], qch))
        out.append(b"\t".join([batch.names[i], b"+" if fw else b"-", rn, str(int(h["toff"])).encode(), seq, qual,
                               str(int(h["oms"])).encode(), b",".join(parts)]) + b"\n")
    return b"".join(out)


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def run_reference(args: list[str], index, reads, extra_env=None) -> tuple[bytes, str]:
    """Run the compiled reference; returns (hit file bytes, stderr)."""
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".out", delete=False) as tf:
        outp = tf.name
    try:
        cmd = [str(REF_ALIGN), *args, "-p", "1", "-x", str(index), *([str(reads)] if not isinstance(reads, list) else reads), outp]
        p = subprocess.run(cmd, capture_output=True, text=True, env=extra_env)
        if p.returncode != 0:
            raise RuntimeError(f"reference failed: {' '.join(cmd)}\n{p.stderr}")
        with open(outp, "rb") as f:
            return f.read(), p.stderr
    finally:
        os.unlink(outp)


# ----------------------------------------------------------------------------------------
# decoding of the kernel's raw output (found / flags / hit records) — mirrors the host-side
# finishRead logic of bowtie_b200 (HitSinkPerThread::finishRead, hit.h:741-786)
# ----------------------------------------------------------------------------------------

BT_HIT_HDR = 5


def decode_device_result(found: np.ndarray, hits: np.ndarray, slots: int, mm_cap: int, pol: Policy) -> AlignResult:
    n = len(found)
    rec_words = BT_HIT_HDR + mm_cap
    hits = hits.reshape(n, slots, rec_words) if n else hits.reshape(0, slots, rec_words)
    nlim = 0xFFFFFFFF if pol.all_hits else pol.khits
    maxed = found > pol.mhits
    nrep = np.where(maxed, 0, np.minimum(found, nlim)).astype(np.uint32)
    out_hits, out_mms = [], []
    for i in np.nonzero(nrep)[0].tolist():
        for s in range(int(nrep[i])):
            rec = hits[i, s]
            w3 = int(rec[3])
            nmm = int(rec[4])
            mm_off = len(out_mms)
            for k in range(min(nmm, mm_cap)):   # reads flagged BT_OVF_MM carry a truncated list
                w = int(rec[BT_HIT_HDR + k])
                out_mms.append((w & 0xFFFF, (w >> 16) & 0xFF, 0))
            out_hits.append((i, int(rec[0]), int(rec[1]), int(rec[2]), w3 & 0xFFFF, (w3 >> 24) & 0xFF, (w3 >> 16) & 0xFF, min(nmm, mm_cap), mm_off))
    h = np.array(out_hits, dtype=HIT_DTYPE) if out_hits else np.zeros(0, HIT_DTYPE)
    m = np.array(out_mms, dtype=MM_DTYPE) if out_mms else np.zeros(0, MM_DTYPE)
    aligned = int(np.count_nonzero(nrep))
    nmaxed = int(np.count_nonzero(maxed))
    ctr = np.array([aligned, n - aligned - nmaxed, nmaxed, int(nrep.sum()), 0], dtype=np.uint64)
    return AlignResult(h, m, nrep, maxed.astype(np.uint8), ctr, None)


def results_equal(a: AlignResult, b: AlignResult) -> tuple[bool, str]:
    if not np.array_equal(a.nhits_per_read, b.nhits_per_read):
        bad = np.nonzero(a.nhits_per_read != b.nhits_per_read)[0]
        return False, f"nhits_per_read differ at reads {bad[:10].tolist()}"
    if not np.array_equal(a.maxed, b.maxed):
        return False, "maxed flags differ"
    if len(a.hits) != len(b.hits):
        return False, f"hit counts differ {len(a.hits)} vs {len(b.hits)}"
    for f in ("read", "tidx", "toff", "oms", "cost", "fw", "stratum", "nmm"):
        if not np.array_equal(a.hits[f], b.hits[f]):
            bad = np.nonzero(a.hits[f] != b.hits[f])[0]
            return False, f"hit field {f} differs at hits {bad[:10].tolist()} (reads {a.hits['read'][bad[:10]].tolist()})"
    for i in range(len(a.hits)):
        ma = a.mms[int(a.hits["mm_off"][i]): int(a.hits["mm_off"][i]) + int(a.hits["nmm"][i])]
        mb = b.mms[int(b.hits["mm_off"][i]): int(b.hits["mm_off"][i]) + int(b.hits["nmm"][i])]
        if not (np.array_equal(ma["pos"], mb["pos"]) and np.array_equal(ma["refc"], mb["refc"])):
            return False, f"mismatch list differs at hit {i} (read {int(a.hits['read'][i])})"
    if not np.array_equal(a.counters, b.counters):
        return False, f"counters differ {a.counters} vs {b.counters}"
    return True, ""


class _DevPolicy(C.Structure):
    _fields_ = [("mode", C.c_int32), ("mms", C.c_int32), ("seedLen", C.c_int32), ("qualThresh", C.c_uint32),
                ("maxBts", C.c_uint32), ("khits", C.c_uint32), ("mhits", C.c_uint32), ("allHits", C.c_int32),
                ("nofw", C.c_int32), ("norc", C.c_int32), ("maqRound", C.c_int32)]


def dev_policy(pol: Policy) -> _DevPolicy:
    return _DevPolicy(pol.mode, pol.mms, pol.seed_len, pol.qual_thresh, pol.max_bts, pol.khits, pol.mhits,
                      int(pol.all_hits), int(pol.nofw), int(pol.norc), int(pol.maq_round))


class HostEmu:
    """Test-only host build of the device state machine (tests/host_emu/emu.cpp)."""

    def __init__(self) -> None:
        ensure_oracle_built()
        d = ROOT / "tests" / "host_emu"
        prof = bool(os.environ.get("BT_EMU_PROFILE"))         # tools/pc_hist.py: loop trip counts of the rare blocks
        so = d / ("libbtemu_prof.so" if prof else "libbtemu.so")
        srcs = [d / "emu.cpp", ROOT / "bowtie_b200" / "csrc" / "bt_core.cuh", ROOT / "bowtie_b200" / "csrc" / "bt_native.cuh", ROOT / "bowtie_b200" / "csrc" / "bt_ctxq.cuh",
                ORACLE_DIR / "bt_oracle.c"]
        def stale():
            return not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs)
        if stale():                                          # several pytest workers may get here at once: one builds, the others wait
            import fcntl
            with open(d / ".emu.lock", "w") as lk:
                fcntl.flock(lk, fcntl.LOCK_EX)
                if stale():
                    tmp = d / f".libbtemu.{os.getpid()}.so"
                    subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared"] + (["-DBT_EMU_PROFILE"] if prof else []) + ["-o", str(tmp), str(d / "emu.cpp"),
                                    str(ORACLE_DIR / "bt_oracle.c")], check=True, capture_output=True)
                    os.replace(tmp, so)
        L = C.CDLL(str(so))
        L.emu_index_load.restype = C.c_void_p
        L.emu_index_load.argtypes = [C.c_char_p, C.c_int]
        L.emu_index_free.argtypes = [C.c_void_p]
        L.emu_lf.restype = C.c_uint32
        L.emu_lf.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.emu_lf_ex.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.emu_row_l.restype = C.c_int
        L.emu_row_l.argtypes = [C.c_void_p, C.c_uint32]
        L.emu_align.restype = C.c_int
        L.emu_align.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_DevPolicy), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.emu_align_sliced.restype = C.c_int
        L.emu_align_sliced.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_DevPolicy), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        self.L = L
        self._idx = {}

    def index(self, base, mirror: bool):
        key = (str(base), int(mirror))
        if key not in self._idx:
            h = self.L.emu_index_load(str(base).encode(), int(mirror))
            if not h:
                raise RuntimeError("emu index load failed")
            self._idx[key] = h
        return self._idx[key]

    def align_sliced(self, base, batch: ReadBatch, pol: Policy, budget0: int, growth: int, slots=None, mm_cap=8, R=None, FCAP=8, PCAP=64,
                     slot_FCAP=16, slot_PCAP=1024):
        """The batch under the kernels' time slicing: reads beyond `budget0` transitions (or with a full seedling list) are suspended into a
        checkpoint slot and resumed with `growth` x the budget (0: unlimited), with a poisoned lane in between.  Returns (result, flags, suspensions)."""
        fw = self.index(base, False)
        bw = self.index(base, True) if (pol.mode == 1 or pol.mms > 0) else None
        n = len(batch)
        if slots is None:
            slots = 64 if pol.all_hits else pol.khits
        maxlen = int((batch.offs[1:] - batch.offs[:-1]).max()) if n else 1
        if R is None:
            R = 8 * maxlen
        found = np.zeros(n, np.uint32)
        flags = np.zeros(n, np.uint32)
        hits = np.zeros(n * slots * (BT_HIT_HDR + mm_cap), np.uint32)
        stats = np.zeros(8, np.uint64)
        nsusp = np.zeros(1, np.uint64)
        cp = dev_policy(pol)
        rc = self.L.emu_align_sliced(fw, bw, C.byref(cp), n, batch.seq_codes.ctypes.data, batch.qual_cat.ctypes.data,
                                     batch.offs.ctypes.data, batch.seeds.ctypes.data, found.ctypes.data, flags.ctypes.data,
                                     hits.ctypes.data, slots, mm_cap, R, FCAP, PCAP, slot_FCAP, slot_PCAP, budget0, growth,
                                     stats.ctypes.data, nsusp.ctypes.data)
        if rc:
            raise RuntimeError(f"emu_align_sliced rc={rc}")
        res = decode_device_result(found, hits, slots, mm_cap, pol)
        res.stats = dict(zip(["lfex", "lf", "chase", "ftab", "offs", "backtracks", "iters", "blockloads"], stats.tolist()))
        self.raw = (found, hits)
        return res, flags, int(nsusp[0])

    def align(self, base, batch: ReadBatch, pol: Policy, slots=None, mm_cap=8, R=None, FCAP=16, PCAP=256):
        fw = self.index(base, False)
        bw = self.index(base, True) if (pol.mode == 1 or pol.mms > 0) else None
        n = len(batch)
        if slots is None:
            slots = 64 if pol.all_hits else pol.khits
        maxlen = int((batch.offs[1:] - batch.offs[:-1]).max()) if n else 1
        if R is None:
            R = 8 * maxlen
        found = np.zeros(n, np.uint32)
        flags = np.zeros(n, np.uint32)
        hits = np.zeros(n * slots * (BT_HIT_HDR + mm_cap), np.uint32)
        stats = np.zeros(8, np.uint64)
        self.iters_per_read = np.zeros(n, np.uint32)
        cp = dev_policy(pol)
        rc = self.L.emu_align(fw, bw, C.byref(cp), n, batch.seq_codes.ctypes.data, batch.qual_cat.ctypes.data,
                              batch.offs.ctypes.data, batch.seeds.ctypes.data, found.ctypes.data, flags.ctypes.data,
                              hits.ctypes.data, slots, mm_cap, R, FCAP, PCAP, stats.ctypes.data, self.iters_per_read.ctypes.data)
        if rc:
            raise RuntimeError(f"emu_align rc={rc}")
        res = decode_device_result(found, hits, slots, mm_cap, pol)
        res.stats = dict(zip(["lfex", "lf", "chase", "ftab", "offs", "backtracks", "iters", "blockloads"], stats.tolist()))
        return res, flags
