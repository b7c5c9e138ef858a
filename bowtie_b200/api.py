"""ctypes binding of libbowtie_b200.so (C ABI: include/bowtie_b200.h)."""
from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
BT_HIT_HDR_WORDS = 5
OVF_STACK, OVF_FRAME, OVF_PART, OVF_HITS, OVF_MM = 1, 2, 4, 8, 16


def lib_path() -> Path:
    """The in-tree CUDA library; BOWTIE_B200_LIB selects a tuning variant built by `make variants` (development only)."""
    import os
    v = os.environ.get("BOWTIE_B200_LIB")
    return Path(v) if v else _PKG / "libbowtie_b200.so"


def build_library(force: bool = False) -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    so = lib_path()
    csrc = _PKG / "csrc"
    srcs = [p for p in csrc.iterdir() if p.suffix in (".cu", ".cuh", ".h")] + list((csrc / "host").glob("*.cpp")) + [_PKG.parent / "include" / "bowtie_b200.h"]
    clis = [_PKG / "bowtie-b200-align", _PKG / "bowtie-b200-build"]
    newest = max(s.stat().st_mtime for s in srcs)

    def stale() -> bool:
        return force or not so.exists() or not all(c.exists() for c in clis) or min([so.stat().st_mtime] + [c.stat().st_mtime for c in clis]) < newest
    if stale():
        import fcntl
        with open(_PKG / ".build.lock", "w") as lk:            # several ranks / pytest workers may get here at once: one builds, the rest wait
            fcntl.flock(lk, fcntl.LOCK_EX)
            if stale():
                p = subprocess.run(["make", "-C", str(_PKG / "csrc")], capture_output=True, text=True)
                if p.returncode != 0:
                    raise RuntimeError("building libbowtie_b200.so failed:\n" + p.stdout + p.stderr)
    return so


class _Policy(C.Structure):
    _fields_ = [("mode", C.c_int32), ("mms", C.c_int32), ("seed_len", C.c_int32), ("qual_thresh", C.c_uint32),
                ("max_bts", C.c_uint32), ("khits", C.c_uint32), ("mhits", C.c_uint32), ("all_hits", C.c_int32),
                ("nofw", C.c_int32), ("norc", C.c_int32), ("maq_round", C.c_int32),
                ("best", C.c_int32), ("strata", C.c_int32), ("max_bts_best", C.c_uint32), ("sample_max", C.c_int32),
                ("paired", C.c_int32), ("min_ins", C.c_uint32), ("max_ins", C.c_uint32), ("mate1fw", C.c_int32), ("mate2fw", C.c_int32),
                ("pair_tries", C.c_uint32)]


class _ReadBatch(C.Structure):
    _fields_ = [("nreads", C.c_uint32), ("seq", C.c_void_p), ("qual", C.c_void_p), ("offs", C.c_void_p),
                ("seeds", C.c_void_p), ("sel", C.c_void_p), ("nsel", C.c_uint32), ("max_len", C.c_uint32)]


class _HitBatch(C.Structure):
    _fields_ = [("found", C.c_void_p), ("flags", C.c_void_p), ("hits", C.c_void_p), ("slots", C.c_uint32),
                ("mm_cap", C.c_uint32)]


class _Info(C.Structure):
    _fields_ = [("len", C.c_uint32), ("n_refs", C.c_uint32), ("off_rate", C.c_int32), ("ftab_chars", C.c_int32),
                ("has_mirror", C.c_int32), ("device_bytes", C.c_uint64)]


class _Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("lfex", "lf", "chase", "ftab", "offs", "backtracks", "iters", "block_loads")]


_LIB = None


def load_library() -> C.CDLL:
    """Load libbowtie_b200.so; raises if it has not been built (no fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = lib_path()
    if not so.exists():
        raise RuntimeError(f"{so} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(bowtie_b200 has no CPU search path)")
    L = C.CDLL(str(so))
    L.bt_abi_version.restype = C.c_int
    L.bt_last_error.restype = C.c_char_p
    L.bt_index_load.restype = C.c_int
    L.bt_index_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.bt_index_free.argtypes = [C.c_void_p]
    L.bt_index_info.argtypes = [C.c_void_p, C.POINTER(_Info)]
    L.bt_index_refname.restype = C.c_char_p
    L.bt_index_refname.argtypes = [C.c_void_p, C.c_uint32]
    L.bt_index_reflen.restype = C.c_uint32
    L.bt_index_reflen.argtypes = [C.c_void_p, C.c_uint32]
    L.bt_policy_init.argtypes = [C.POINTER(_Policy)]
    for fn in (L.bt_align_batch, L.bt_align_batch_device, L.bt_context_align, L.bt_context_align_async, L.bt_context_align_device):
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(_Policy), C.POINTER(_ReadBatch), C.POINTER(_HitBatch), C.c_void_p]
    L.bt_context_create.restype = C.c_int
    L.bt_context_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    L.bt_context_free.argtypes = [C.c_void_p]
    for fn in (L.bt_context_sync, L.bt_context_join):
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p]
    L.bt_stats_get.argtypes = [C.c_void_p, C.POINTER(_Stats), C.c_int]
    L.bt_debug_lf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
    L.bt_index_build.restype = C.c_int
    L.bt_index_build.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.bt_index_build_text.restype = C.c_int
    L.bt_index_build_text.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.c_uint32, C.c_char_p, C.c_int, C.c_int, C.c_int]
    _LIB = L
    return L


def build_index(fasta, out_base, off_rate: int = 5, ftab_chars: int = 10, device: int = 0) -> None:
    """bowtie-build on the GPU (bt_index_build): writes out_base.{1,2,3,4}.ebwt and out_base.rev.{1,2}.ebwt."""
    L = load_library()
    files = [str(f).encode() for f in ([fasta] if isinstance(fasta, (str, Path)) else fasta)]
    arr = (C.c_char_p * len(files))(*files)
    if L.bt_index_build(arr, len(files), str(out_base).encode(), int(off_rate), int(ftab_chars), int(device)) != 0:
        raise RuntimeError("bt_index_build: " + L.bt_last_error().decode())


def build_index_text(codes: np.ndarray, recs, names, out_base, off_rate: int = 5, ftab_chars: int = 10, device: int = 0) -> None:
    """bt_index_build_text: the index of an in-memory reference.  codes: uint8 base codes 0..3 of the joined unambiguous
    characters; recs: (off, len, first) per record (RefRecord, ref_read.h:57-88); names: one per sequence."""
    L = load_library()
    codes = np.ascontiguousarray(codes, np.uint8)
    r = np.ascontiguousarray(np.asarray(recs, np.uint32).reshape(-1, 3))
    nm = [str(n).encode() for n in names]
    arr = (C.c_char_p * len(nm))(*nm)
    if L.bt_index_build_text(codes.ctypes.data, codes.size, r.ctypes.data, r.shape[0], arr, len(nm), str(out_base).encode(),
                             int(off_rate), int(ftab_chars), int(device)) != 0:
        raise RuntimeError("bt_index_build_text: " + L.bt_last_error().decode())


@dataclass
class Policy:
    """Search policy; field meanings follow the reference's options (ebwt_search.cpp:153-253)."""
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
    best: bool = False       # --best: the reference's best-first ("stateful") aligners; implied by strata and -v 3
    strata: bool = False     # --strata
    max_bts_best: int = 800  # --maxbts on the best-first path
    sample_max: bool = False # -M: keep every hit up to the mhits ceiling
    paired: bool = False     # reads 2p, 2p+1 are the mates of pair p
    min_ins: int = 0         # -I
    max_ins: int = 250       # -X
    mate1fw: bool = True     # --fr
    mate2fw: bool = False
    pair_tries: int = 100    # --pairtries

    def to_c(self) -> _Policy:
        return _Policy(self.mode, self.mms, self.seed_len, self.qual_thresh, self.max_bts, self.khits, self.mhits,
                       int(self.all_hits), int(self.nofw), int(self.norc), int(self.maq_round),
                       int(self.best or self.strata or self.sample_max), int(self.strata), self.max_bts_best, int(self.sample_max),
                       int(self.paired), self.min_ins, self.max_ins, int(self.mate1fw), int(self.mate2fw), self.pair_tries)

    @property
    def stateful(self) -> bool:
        return self.best or self.strata or self.sample_max or (self.mode == 0 and self.mms == 3)

    @property
    def needs_mirror(self) -> bool:
        return self.mode == 1 or self.mms > 0

    @property
    def report_limit(self) -> int:
        return 0xFFFFFFFF if self.all_hits else self.khits


@dataclass
class Stats:
    lfex: int
    lf: int
    chase: int
    ftab: int
    offs: int
    backtracks: int
    iters: int
    block_loads: int

    @property
    def side_fetches(self) -> int:
        """mapLFEx = 2 side fetches, every other LF = 1."""
        return 2 * self.lfex + self.lf

    @property
    def algorithmic_bytes(self) -> int:
        """bytes = 64·N_side + 4·N_offs + 8·N_ftab."""
        return 64 * self.side_fetches + 4 * self.offs + 8 * self.ftab


class Index:
    """Device-resident forward (+ mirror) index loaded from unmodified .ebwt files."""

    def __init__(self, basename: str, need_mirror: bool = True, device: int = 0) -> None:
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.bt_index_load(str(basename).encode(), int(need_mirror), int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("bt_index_load: " + self.L.bt_last_error().decode())
        self.h = h
        info = _Info()
        self.L.bt_index_info(self.h, C.byref(info))
        self.len, self.n_refs, self.off_rate, self.ftab_chars = info.len, info.n_refs, info.off_rate, info.ftab_chars
        self.has_mirror, self.device_bytes, self.device = bool(info.has_mirror), int(info.device_bytes), device
        self.refnames = []
        i = 0
        while True:
            nm = self.L.bt_index_refname(self.h, i)
            if nm is None:
                break
            self.refnames.append(nm)
            i += 1
        self.reflens = [int(self.L.bt_index_reflen(self.h, i)) for i in range(self.n_refs)]

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.bt_index_free(self.h)
            self.h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what}: " + self.L.bt_last_error().decode())

    def align(self, seq: np.ndarray, qual: np.ndarray, offs: np.ndarray, seeds: np.ndarray, pol: Policy,
              slots: int | None = None, mm_cap: int = 8, sel: np.ndarray | None = None, out=None, stream: int = 0):
        """Host buffers in, host buffers out (bt_align_batch).  Returns (found, flags, hits[n, slots, 5+mm_cap])."""
        n = len(seeds)
        if slots is None:
            slots = 16 if pol.all_hits else pol.khits
        seq = np.ascontiguousarray(seq, np.uint8); qual = np.ascontiguousarray(qual, np.uint8)
        offs = np.ascontiguousarray(offs, np.uint64); seeds = np.ascontiguousarray(seeds, np.uint32)
        rw = BT_HIT_HDR_WORDS + mm_cap
        if out is None:
            found = np.zeros(n, np.uint32); flags = np.zeros(n, np.uint32); hits = np.zeros((n, slots, rw), np.uint32)
        else:
            found, flags, hits = out
        rb = _ReadBatch(n, seq.ctypes.data, qual.ctypes.data, offs.ctypes.data, seeds.ctypes.data, None, 0, 0)
        if sel is not None:
            sel = np.ascontiguousarray(sel, np.uint32)
            rb.sel = sel.ctypes.data
            rb.nsel = len(sel)
        hb = _HitBatch(found.ctypes.data, flags.ctypes.data, hits.ctypes.data, slots, mm_cap)
        cp = pol.to_c()
        self._check(self.L.bt_align_batch(self.h, C.byref(cp), C.byref(rb), C.byref(hb), C.c_void_p(stream)), "bt_align_batch")
        return found, flags, hits

    def align_device(self, seq_ptr: int, qual_ptr: int, offs_ptr: int, seeds_ptr: int, n: int, max_len: int, pol: Policy,
                     found_ptr: int, flags_ptr: int, hits_ptr: int, slots: int, mm_cap: int, stream: int = 0) -> None:
        """Device pointers in and out; enqueues on `stream` without synchronising (bt_align_batch_device)."""
        rb = _ReadBatch(n, seq_ptr, qual_ptr, offs_ptr, seeds_ptr, None, 0, max_len)
        hb = _HitBatch(found_ptr, flags_ptr, hits_ptr, slots, mm_cap)
        cp = pol.to_c()
        self._check(self.L.bt_align_batch_device(self.h, C.byref(cp), C.byref(rb), C.byref(hb), C.c_void_p(stream)),
                    "bt_align_batch_device")

    def stats(self, reset: bool = False) -> Stats:
        s = _Stats()
        self._check(self.L.bt_stats_get(self.h, C.byref(s), int(reset)), "bt_stats_get")
        return Stats(*(int(getattr(s, k)) for k, _ in _Stats._fields_))

    def debug_lf(self, rows: np.ndarray, mirror: bool = False) -> np.ndarray:
        rows = np.ascontiguousarray(rows, np.uint32)
        out = np.zeros((len(rows), 5), np.uint32)
        self._check(self.L.bt_debug_lf(self.h, int(mirror), rows.ctypes.data, len(rows), out.ctypes.data), "bt_debug_lf")
        return out


class Context:
    """One in-flight batch over a shared Index (bt_context_t): scratch + staging.  Use one per CUDA stream
    to keep several batches in flight (the reference's equivalent: one worker thread per `-p`)."""

    def __init__(self, index: Index) -> None:
        self.ix, self.L = index, index.L
        h = C.c_void_p()
        index._check(self.L.bt_context_create(index.h, C.byref(h)), "bt_context_create")
        self.h = h

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.bt_context_free(self.h)
            self.h = None

    def align_device(self, seq_ptr, qual_ptr, offs_ptr, seeds_ptr, n, max_len, pol: Policy, found_ptr, flags_ptr, hits_ptr,
                     slots, mm_cap, stream=0) -> None:
        rb = _ReadBatch(n, seq_ptr, qual_ptr, offs_ptr, seeds_ptr, None, 0, max_len)
        hb = _HitBatch(found_ptr, flags_ptr, hits_ptr, slots, mm_cap)
        cp = pol.to_c()
        self.ix._check(self.L.bt_context_align_device(self.h, C.byref(cp), C.byref(rb), C.byref(hb), C.c_void_p(stream)), "bt_context_align_device")

    def align_async(self, seq, qual, offs, seeds, pol: Policy, out, slots, mm_cap, stream=0) -> None:
        """Pinned host arrays in/out; enqueue H2D + kernels + D2H on `stream` and return (bt_context_align_async)."""
        found, flags, hits = out
        rb = _ReadBatch(len(seeds), seq.ctypes.data, qual.ctypes.data, offs.ctypes.data, seeds.ctypes.data, None, 0, 0)
        hb = _HitBatch(found.ctypes.data, flags.ctypes.data, hits.ctypes.data, slots, mm_cap)
        cp = pol.to_c()
        self.ix._check(self.L.bt_context_align_async(self.h, C.byref(cp), C.byref(rb), C.byref(hb), C.c_void_p(stream)), "bt_context_align_async")

    def sync(self, stream=0) -> None:
        self.ix._check(self.L.bt_context_sync(self.h, C.c_void_p(stream)), "bt_context_sync")

    def join(self, stream=0) -> None:
        """Make `stream` wait for this context's outstanding batch, including its side-stream passes."""
        self.ix._check(self.L.bt_context_join(self.h, C.c_void_p(stream)), "bt_context_join")


def decode_hits(found: np.ndarray, hits: np.ndarray, pol: Policy):
    """Apply HitSinkPerThread::finishRead (hit.h:741-786) to the raw kernel output.

    Returns (nreported[n], maxed[n]) — read i reports the first nreported[i] records of hits[i]."""
    maxed = found > np.uint32(pol.mhits) if pol.mhits != 0xFFFFFFFF else np.zeros(len(found), bool)
    nrep = np.where(maxed, 0, np.minimum(found, pol.report_limit)).astype(np.uint32)
    return nrep, maxed
