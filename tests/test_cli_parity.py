"""Drop-in check of the bowtie-compatible host driver (bowtie_b200/bowtie-b200-align): its hit file and
stderr summary must equal the unmodified reference binary's (oracle/_ref/bowtie-align-s) byte for byte.

* not-gpu variant: the driver runs against tests/host_emu/shim/libbowtie_b200.so (the device state machine
  compiled for the host, test-only) — this checks the HOST code: option parsing, read parsing, formatting.
* gpu variant: the same comparisons with the real CUDA library.
"""
import os
import subprocess
from pathlib import Path

import pytest

from helpers import FIXTURES, REF_ALIGN, ROOT, ensure_oracle_built, have_reference

CLI = ROOT / "bowtie_b200" / "bowtie-b200-align"
SHIM_DIR = ROOT / "tests" / "host_emu" / "shim"

CASES = [
    ("n2-default", ["-n", "2"], "fq"),
    ("v0", ["-v", "0"], "fq"),
    ("v2-k3", ["-v", "2", "-k", "3"], "fq"),
    ("n2-a-cost", ["-n", "2", "-a", "--cost"], "fq"),
    ("n3-m2", ["-n", "3", "-m", "2"], "fq"),
    ("n2-sam", ["-n", "2", "-S"], "fq"),
    ("v2-sam-a", ["-v", "2", "-a", "-S", "--mapq", "42", "--sam-RG", "ID:x", "--sam-RG", "SM:y"], "fq"),
    ("n2-sam-m1-nounal", ["-n", "2", "-m", "1", "-S", "--no-unal"], "fq"),
    ("n2-sam-nohead", ["-n", "2", "-S", "--sam-nohead"], "fq"),
    ("v1-fasta", ["-v", "1", "-f"], "fa"),
    ("n2-raw", ["-n", "2", "-r"], "raw"),
    ("n2-trim", ["-n", "2", "-5", "3", "-3", "2"], "fq"),
    ("n1-skip-upto", ["-n", "1", "-s", "100", "-u", "500"], "fq"),
    ("n2-suppress-offbase", ["-n", "2", "--suppress", "1,5,6", "-B", "1"], "fq"),
    ("n2-refidx", ["-n", "2", "--refidx"], "fq"),
    ("n2-fullref-nofw", ["-n", "2", "--fullref", "--nofw"], "fq"),
    ("n2-l20-e100-nomaqround", ["-n", "2", "-l", "20", "-e", "100", "--nomaqround"], "fq"),
    ("v2-cmdline", ["-a", "-v", "2", "--suppress", "1,5,6,7", "-c"], "cmd"),     # MANUAL.markdown:246-253 Example 1
    ("n2-seed7", ["-n", "2", "--seed", "7", "-k", "2"], "fq"),
    # the best-first ("stateful") path: SURVEY.md §8(c) goldens and the options that select it
    ("best-n2", ["-n", "2", "--best"], "fq"),                                   # md5 89343691214487f3c47e35cbe4cf5a15
    ("best-n2-strata-k5", ["-n", "2", "--best", "--strata", "-k", "5"], "fq"),  # md5 fa7667c8d4e433be247da97e79a61cbb
    ("v3", ["-v", "3"], "fq"),                                                  # md5 13e0579a9fc872cede00a27d9bd9c5bf
    ("v3-a-strata", ["-v", "3", "-a", "--best", "--strata"], "fq"),
    ("best-v2-k3", ["-v", "2", "--best", "-k", "3"], "fq"),
    ("best-v1-strata-m3", ["-v", "1", "--best", "--strata", "-m", "3", "-a"], "fq"),
    ("best-v0", ["-v", "0", "--best"], "fq"),
    ("best-n0", ["-n", "0", "--best", "--strata", "-k", "2"], "fq"),
    ("best-n1-a", ["-n", "1", "--best", "-a"], "fq"),
    ("best-n3-strata-a-nomaqround", ["-n", "3", "--best", "--strata", "-a", "--nomaqround"], "fq"),
    ("best-n2-m1", ["-n", "2", "--best", "-m", "1"], "fq"),
    ("M1", ["-n", "2", "-M", "1"], "fq"),
    ("M2-strata-sam", ["-n", "2", "-M", "2", "--best", "--strata", "-k", "2", "-S"], "fq"),
    ("best-n2-sam", ["-n", "2", "--best", "-S"], "fq"),
    ("best-n2-norc-l20", ["-n", "2", "--best", "--norc", "-l", "20", "-e", "100"], "fq"),
    ("best-n3-maxbts", ["-n", "3", "--best", "--maxbts", "20", "-a", "--strata"], "fq"),
    ("best-n2-tryhard", ["-n", "2", "--best", "-y", "-a"], "fq"),
    ("best-fasta", ["-v", "3", "-f", "-k", "2"], "fa"),
]

GOLDEN_MD5 = {"best-n2": "89343691214487f3c47e35cbe4cf5a15", "best-n2-strata-k5": "fa7667c8d4e433be247da97e79a61cbb", "v3": "13e0579a9fc872cede00a27d9bd9c5bf"}


@pytest.fixture(scope="module")
def cli():
    ensure_oracle_built()
    if not have_reference():
        pytest.skip("reference binary / fixtures not available")
    import bowtie_b200
    bowtie_b200.build_library()            # also builds the driver
    if not CLI.exists():
        pytest.fail("bowtie-b200-align was not built")
    return CLI


def build_shim():
    """The emulation shim, rebuilt when a source changed; several pytest workers may get here at once: one builds (file lock,
    atomic rename), the others wait."""
    import fcntl
    so = SHIM_DIR / "libbowtie_b200.so"
    csrc = ROOT / "bowtie_b200" / "csrc"
    srcs = [ROOT / "tests" / "host_emu" / "abi_shim.cpp", csrc / "bt_core.cuh", ROOT / "oracle" / "bt_oracle.c", csrc / "bt_best.cuh", csrc / "bt_best_prog.h",
            csrc / "bt_build.h", csrc / "bt_build_sa.cuh", csrc / "bt_io.cuh", csrc / "bt_io_run.h", csrc / "bt_prog.h", csrc / "bt_native.cuh",
            ROOT / "tests" / "host_emu" / "bsa_host.h", ROOT / "include" / "bowtie_b200.h"]

    def stale():
        return not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs)
    if stale():
        SHIM_DIR.mkdir(exist_ok=True)
        with open(SHIM_DIR / ".lock", "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            if stale():
                tmp = SHIM_DIR / f".libbowtie_b200.{os.getpid()}.so"
                subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(tmp), str(srcs[0]), str(srcs[2]), "-lz"], check=True, capture_output=True)
                os.replace(tmp, so)
    return so


def reads_arg(kind):
    return {"fq": str(FIXTURES / "e_coli_1000.fq"), "fa": str(FIXTURES / "e_coli_1000.fa"), "raw": str(FIXTURES / "e_coli_1000.raw"),
            "cmd": "ATGCATCATGCGCCAT"}[kind]


def run(exe, flags, kind, out, env=None, ref=False):
    cmd = [str(exe), *flags] + (["-p", "1"] if ref else []) + ["-x", str(FIXTURES / "e_coli"), reads_arg(kind), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    body = Path(out).read_bytes()
    body = b"".join(l for l in body.splitlines(keepends=True) if not l.startswith(b"@PG"))   # @PG embeds the command line
    summary = "\n".join(l for l in p.stderr.splitlines() if l.startswith("#") or l.startswith("Reported") or l.startswith("No alignments"))
    return body, summary


def compare(cli, flags, kind, tmp_path, env):
    ref_body, ref_sum = run(REF_ALIGN, flags, kind, tmp_path / "ref.out", ref=True)
    our_body, our_sum = run(cli, flags, kind, tmp_path / "our.out", env=env)
    assert our_body == ref_body
    assert our_sum == ref_sum
    assert len(ref_body) > 0
    return our_body


@pytest.mark.parametrize("name,flags,kind", CASES, ids=[c[0] for c in CASES])
def test_cli_host_logic_matches_reference(name, flags, kind, cli, tmp_path):
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    body = compare(cli, flags, kind, tmp_path, env)
    if name in GOLDEN_MD5:
        import hashlib
        assert hashlib.md5(body).hexdigest() == GOLDEN_MD5[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name,flags,kind", CASES, ids=[c[0] for c in CASES])
def test_cli_gpu_matches_reference(name, flags, kind, cli, tmp_path):
    env = {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"}
    body = compare(cli, flags, kind, tmp_path, env)
    if name in GOLDEN_MD5:
        import hashlib
        assert hashlib.md5(body).hexdigest() == GOLDEN_MD5[name]


def test_cli_rejects_unsupported_combinations(cli, tmp_path):
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    # (never hand a fixture to the driver as a trailing positional argument: with paired input it is the OUTPUT file)
    p = subprocess.run([str(cli), "--large-index", "-x", str(FIXTURES / "e_coli"), "-c", "ACGTACGTACGT"], capture_output=True, text=True, env=env)
    assert p.returncode != 0 and "large" in p.stderr
    p = subprocess.run([str(cli), "--no-such-option", "-x", str(FIXTURES / "e_coli"), "-c", "ACGTACGTACGT"], capture_output=True, text=True, env=env)
    assert p.returncode != 0
    # ebwt_search.cpp:883-890
    p = subprocess.run([str(cli), "--strata", "-x", str(FIXTURES / "e_coli"), "-c", "ACGTACGTACGT"], capture_output=True, text=True, env=env)
    assert p.returncode != 0 and "--strata must be combined with --best" in p.stderr
    p = subprocess.run([str(cli), "--best", "--strata", "-x", str(FIXTURES / "e_coli"), "-c", "ACGTACGTACGT"], capture_output=True, text=True, env=env)
    assert p.returncode != 0 and "--strata has no effect" in p.stderr


def check_dumps(cli, tmp_path, env):
    """--al / --un / --max (HitSink::dumpAlign/dumpUnal/dumpMaxed, hit.h:385-492): the reads, as they stood in the input, split by
    outcome; pairs go to <base>_1/<base>_2; maxed reads fall back to --un without --max."""
    fq, fa, raw = FIXTURES / "e_coli_1000.fq", FIXTURES / "e_coli_1000.fa", FIXTURES / "e_coli_1000.raw"
    m1, m2 = FIXTURES / "e_coli_1000_1.fq", FIXTURES / "e_coli_1000_2.fq"
    runs = [(["-n", "2", "-m", "1"], ["-q", str(fq)], True), (["-n", "2", "-M", "1"], ["-q", str(fq)], True), (["-v", "1", "-m", "1"], ["-q", str(fq)], False),
            (["-n", "2", "-m", "1"], ["-f", str(fa)], True), (["-n", "2", "-m", "1"], ["-r", str(raw)], True),
            (["-n", "2", "-m", "1", "-X", "150"], ["-1", str(m1), "-2", str(m2)], True)]
    for flags, src, with_max in runs:
        got = []
        for tag, exe, e, extra in (("ref", REF_ALIGN, None, ["-p", "1"]), ("our", cli, env, [])):
            d = tmp_path / f"{tag}_{len(got)}_{abs(hash(tuple(flags + src))) % 10000}"
            d.mkdir()
            dump = ["--un", str(d / "un.fq"), "--al", str(d / "al.fq")] + (["--max", str(d / "max.fq")] if with_max else [])
            p = subprocess.run([str(exe), *flags, *extra, *dump, "-x", str(FIXTURES / "e_coli"), *src, str(d / "hits.out")], capture_output=True, text=True, env=e)
            assert p.returncode == 0, p.stderr
            got.append({f.name: f.read_bytes() for f in sorted(d.iterdir()) if f.name != "hits.out"})
        assert got[0].keys() == got[1].keys(), (flags, src)
        assert got[0] == got[1], (flags, src)
        assert any(len(v) for v in got[0].values())


def test_cli_read_dumps(cli, tmp_path):
    build_shim()
    check_dumps(cli, tmp_path, dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR)))


@pytest.mark.gpu
def test_cli_read_dumps_gpu(cli, tmp_path):
    check_dumps(cli, tmp_path, {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"})


def odd_fastq(path, n=9000):
    """The e_coli reads repeated to `n` records (enough for the multi-threaded FASTQ path to split the work) with the oddities
    FastqPatternSource tolerates (pat.cpp:797-975) sprinkled in: CRLF records, '.', '-', lower case,
    an empty name, and a last record without a newline."""
    import random
    rng = random.Random(5)
    lines = (FIXTURES / "e_coli_1000.fq").read_text().splitlines()
    recs = [lines[i:i + 4] for i in range(0, len(lines), 4)]
    out = []
    for k in range(n):
        name, seq, plus, qual = recs[k % len(recs)]
        name = f"{name}_{k}"
        what = rng.randrange(60)
        if what == 0:
            out.append("\r\n".join([name, seq, plus, qual]) + "\r\n")
            continue
        if what == 2:
            seq = seq[:5] + "." + seq[6:]
        elif what == 3:
            seq = seq[:7] + "-" + seq[7:]
        elif what == 4:
            seq = seq.lower()
        elif what == 5:
            name = "@"
        out.append("\n".join([name, seq, plus, qual]) + "\n")
    out[-1] = out[-1].rstrip("\n")
    Path(path).write_text("".join(out))


@pytest.mark.parametrize("threads", ["1", "4"])
def test_cli_fastq_fast_path_matches_reference(threads, cli, tmp_path):
    """Reader::fast_batch (whole buffers of well-formed records, parsed by several threads) hands anything unusual to the
    record-at-a-time parser; names, sequences, qualities, per-read seeds and the order must come out as the reference's."""
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    fq = tmp_path / "odd.fq"
    odd_fastq(fq)
    outs = []
    for exe, extra, e in ((REF_ALIGN, ["-p", "1"], None), (cli, ["-p", threads, "--reads-per-batch", "4096"], env)):
        out = tmp_path / f"{Path(exe).name}.out"
        p = subprocess.run([str(exe), "-n", "2", "-k", "2", *extra, "-x", str(FIXTURES / "e_coli"), str(fq), str(out)], capture_output=True, text=True, env=e)
        assert p.returncode == 0, p.stderr
        outs.append((out.read_bytes(), [l for l in p.stderr.splitlines() if l.startswith("#") or l.startswith("Reported")]))
    assert outs[0] == outs[1]
    assert outs[0][0].count(b"\n") > 5000


@pytest.mark.parametrize("flags", [[], ["-S"], ["--best", "-k", "3", "-S"]], ids=["default", "sam", "best-sam"])
def test_cli_parallel_formatting_matches_reference(flags, cli, tmp_path):
    """Batches of >= 16384 units are formatted by several threads into private buffers that are written in read order."""
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    lines = (FIXTURES / "e_coli_1000.fq").read_text().splitlines()
    recs = [lines[i:i + 4] for i in range(0, len(lines), 4)]
    fq = tmp_path / "big.fq"
    fq.write_text("".join(f"{r[0]}_{k}\n{r[1]}\n+\n{r[3]}\n" for k in range(20) for r in recs))
    outs = []
    for exe, extra, e in ((REF_ALIGN, ["-p", "1"], None), (cli, ["-p", "4", "--reads-per-batch", "32768"], env)):
        out = tmp_path / f"{Path(exe).name}.out"
        p = subprocess.run([str(exe), "-n", "2", *flags, *extra, "-x", str(FIXTURES / "e_coli"), str(fq), str(out)], capture_output=True, text=True, env=e)
        assert p.returncode == 0, p.stderr
        outs.append(b"".join(l for l in out.read_bytes().splitlines(keepends=True) if not l.startswith(b"@PG")))
    assert outs[0] == outs[1] and outs[0].count(b"\n") > 13000


@pytest.mark.parametrize("rb", ["1", "7"])
def test_cli_tiny_batches_through_the_pipeline(rb, cli, tmp_path):
    """Hundreds of batches through the parser-thread ring: batch boundaries must not show in the output."""
    import hashlib
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    out = tmp_path / "o.out"
    p = subprocess.run([str(cli), "-n", "2", "-p", "3", "--reads-per-batch", rb, "-x", str(FIXTURES / "e_coli"), str(FIXTURES / "e_coli_1000.fq"), str(out)],
                       capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    assert hashlib.md5(out.read_bytes()).hexdigest() == "7238b0f529dfcdf602f82ae1a754a1bd"      # -n 2
