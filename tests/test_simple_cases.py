"""Known-answer cases transcribed from the reference's own black-box suite (scripts/test/simple_tests.pl:119-900;
the unpaired cases and the -1/-2 paired cases): tiny literal references + reads + the expected {offset => count} per read.
Each case is run through (1) the unmodified reference binary, whose output must contain exactly the expected
offsets (this pins the transcription), and (2) our bowtie-compatible driver, whose hit file and exit status
must equal the reference's byte for byte (default and SAM output).  On the CPU test box the driver runs against the
host emulation of the device state machine (tests/host_emu/shim); under `-m gpu` against the CUDA library.
Edge cases covered: 8-19 bp references (one rank block, ftab entries that point into eftab), CRLF line ends,
blank lines, reads trimmed to nothing, empty reads, -s/-u, qualities too short/long (must abort), -m, edits."""
import os
import subprocess
from pathlib import Path

import pytest

from helpers import REF_ALIGN, REF_BUILD, ROOT, ensure_oracle_built

CLI = ROOT / "bowtie_b200" / "bowtie-b200-align"
SHIM_DIR = ROOT / "tests" / "host_emu" / "shim"
REF19 = "AGCATCGATCAGTATCTGA"

# (name, ref, kind, reads text, extra args list, expected hits per read or None when the run must abort)
C = []
def case(name, ref, kind, text, args=(), hits=None, abort=False):
    C.append((name, ref, kind, text, list(args), None if abort else hits))

case("Cline 1", REF19, "c", "CATCGATCAGTATCTG", hits=[{2: 1}])
case("Cline 2", REF19, "c", "CATCGATCAGTATCTG:IIIIIIIIIIIIIIII", hits=[{2: 1}])
case("Cline 3", REF19, "c", "CATCGATCAGTATCTG:ABCDEDGHIJKLMNOP", hits=[{2: 1}])
case("Cline 7", REF19, "c", "CATCGATCAGTATCTG:IIIIIIIIIIIIIIII", ["--trim3", "4", "--norc"], hits=[{2: 1}])
case("Cline 8", REF19, "c", "CATCGATCAGTATCTG:IIIIIIIIIIIIIIII", ["--trim5", "16"], hits=[{}])
case("Cline 9", REF19, "c", "CATCGATCAGTATCTG:IIIIIIIIIIIIIIII", ["-s", "1"], hits=[])
case("Cline multiread 1", REF19, "c", "CATCGATCAGTATCTG:IIIIIIIIIIIIIIII,ATCGATCAGTATCTG:IIIIIIIIIIIIIII", hits=[{2: 1}, {3: 1}])
case("Cline multiread 2", REF19, "c", "CATCGATCAGTATCTG:IIIIIIIIIIIIIIII,ATCGATCAGTATCTG:IIIIIIIIIIIIIII", ["-u", "1"], hits=[{2: 1}])
case("Fastq 1", REF19, "q", "@r0\nCATCGATCAGTATCTG\n+\nIIIIIIIIIIIIIIII", hits=[{2: 1}])
case("Fastq 2", REF19, "q", "@r0\nCATCGATCAGTATCTG\n+\nIIIIIIIIIIIIIIII\n", hits=[{2: 1}])
case("Fastq 3", REF19, "q", "@r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIIII\n", hits=[{2: 1}])
case("Fastq 4", REF19, "q", "@r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIII\n", abort=True)
case("Fastq 6", REF19, "q", "r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIII\n", abort=True)
case("Fastq 7", REF19, "q", "@r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIIII\n", ["--trim3", "4", "--norc"], hits=[{2: 1}])
case("Fastq 8", REF19, "q", "@r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIIII\n", ["--trim5", "16"], hits=[{}])
case("Fastq 9", REF19, "q", "@r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIIII\n", ["-s", "1"], hits=[])
case("Fastq multiread 1", REF19, "q", "@r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIIII\n@r1\nATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIII\n", hits=[{2: 1}, {3: 1}])
case("Fastq multiread 2", REF19, "q", "@r0\nCATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIIII\n@r1\nATCGATCAGTATCTG\r\n+\nIIIIIIIIIIIIIII\n", ["-u", "1"], hits=[{2: 1}])
# the perl suite feeds this text as *tabbed* input; as FASTQ the reference aborts on the empty quality line (pat.cpp:926)
case("Fastq empty 2 (as FASTQ)", REF19, "q", "\n\n\r\n@r0\nCATCGATCAGTATCTG\n+\nIIIIIIIIIIIIIIII\n@r1\n\n+\n\n@r2\nCATCGATCAGTATCTG\n+\nIIIIIIIIIIIIIIII", abort=True)
case("Fasta 1", REF19, "f", ">r0\nCATCGATCAGTATCTG", hits=[{2: 1}])
case("Fasta 3", REF19, "f", "\n\n\r\n>r0\nCATCGATCAGTATCTG\r\n\n", hits=[{2: 1}])
case("Fasta 6", REF19, "f", "r0\nCATCGATCAGTATCTG\r", abort=True)
case("Fasta 7", REF19, "f", "\n\n\r\n>r0\nCATCGATCAGTATCTG\r\n", ["--trim3", "4", "--norc"], hits=[{2: 1}])
case("Fasta 8", REF19, "f", "\n\n\r\n>r0\nCATCGATCAGTATCTG\r\n", ["--trim3", "16"], hits=[{}])
case("Fasta multiread 1", REF19, "f", "\n\n\r\n>r0\nCATCGATCAGTATCTG\n\n\n\n\r\n>r1\nATCGATCAGTATCTG\n\n", hits=[{2: 1}, {3: 1}])
case("Fasta multiread 2", REF19, "f", "\n\n\r\n>r0\nCATCGATCAGTATCTG\r\n\n\n\r\n>r1\nATCGATCAGTATCTG\r\n", ["-u", "1"], hits=[{2: 1}])
case("Raw 1", REF19, "r", "CATCGATCAGTATCTG", hits=[{2: 1}])
case("Raw 3", REF19, "r", "\n\n\nCATCGATCAGTATCTG\n\n", hits=[{2: 1}])
case("Raw 7", REF19, "r", "\n\n\r\nCATCGATCAGTATCTG\r\n", ["--trim3", "4", "--norc"], hits=[{2: 1}])
case("Raw 8", REF19, "r", "\n\n\r\nCATCGATCAGTATCTG\r\n", ["--trim3", "16"], hits=[{}])
case("Raw multiread 1", REF19, "r", "\n\n\r\nCATCGATCAGTATCTG\n\n\n\n\r\nATCGATCAGTATCTG\n\n", hits=[{2: 1}, {3: 1}])
for mode in (["-v", "0"], ["-n", "0"]):
    case("Checking -m, 1 " + " ".join(mode), "TTGTTCGTTTGTTCGT", "c", "TTGTTCGT", mode + ["-m", "2", "-a"], hits=[{0: 1, 8: 1}])
    case("Checking -m, 2 " + " ".join(mode), "TTGTTCGTTTGTTCGTTTGTTCGT", "c", "TTGTTCGT", mode + ["-m", "2", "-a"], hits=[{}])
    case("Checking edits 3 " + " ".join(mode), "ACGTTCGT", "c", "GTTC", mode, hits=[{2: 1}])
for mode in (["-v", "2"], ["-n", "2"]):
    case("Checking edits 1 " + " ".join(mode), "TTGCCCGT", "c", "TTGTTCGT", mode, hits=[{0: 1}])
    case("Checking edits 2 " + " ".join(mode), "TTGTTCGT", "c", "ACGGGCAA", mode, hits=[{0: 1}])

# paired-end cases (simple_tests.pl:187-234, 319-367, 452-497, 565-610): (name, ref, kind, mate-1 text, mate-2 text, args, expected "lo,hi" per pair)
REFP = "AGCATCGATCAAAAACTGA"
P = []
def pcase(name, kind, t1, t2, args=(), pairhits=()):
    P.append((name, REFP, kind, t1, t2, list(args), list(pairhits)))

pcase("Cline paired 1", "c", "AGCATCGATC:IIIIIIIIII,TCAGTTTTTGA", "TCAGTTTTTGA,AGCATCGATC:IIIIIIIIII", pairhits=[(0, 8), (0, 8)])
pcase("Cline paired 2", "c", "AGCATCGATC:IIIIIIIIII,TCAGTTTTTGA:IIIIIIIIIII", "TCAGTTTTTGA:IIIIIIIIIII,AGCATCGATC:IIIIIIIIII", ["-s", "1"], [(0, 8)])
pcase("Cline paired 3", "c", "AGCATCGATC:IIIIIIIIII,TCAGTTTTTGA:IIIIIIIIIII", "TCAGTTTTTGA:IIIIIIIIIII,AGCATCGATC:IIIIIIIIII", ["-u", "1"], [(0, 8)])
pcase("Cline paired 4", "c", "AGCATCG:IIIIIII", "GATCAAAAACTGA:IIIIIIIIIIIII", ["-3", "7"], [])
pcase("Fastq paired 1", "q", "@r0\nAGCATCGATC\r\n+\nIIIIIIIIII\n@r1\nTCAGTTTTTGA\r\n+\nIIIIIIIIIII\n", "@r0\nTCAGTTTTTGA\n+\nIIIIIIIIIII\n@r1\nAGCATCGATC\r\n+\nIIIIIIIIII", pairhits=[(0, 8), (0, 8)])
pcase("Fastq paired 2", "q", "@r0\nAGCATCGATC\r\n+\nIIIIIIIIII\n@r1\nTCAGTTTTTGA\n+\nIIIIIIIIIII\n", "@r0\nTCAGTTTTTGA\n+\nIIIIIIIIIII\n@r1\nAGCATCGATC\r\n+\nIIIIIIIIII", ["-s", "1"], [(0, 8)])
pcase("Fastq paired 3", "q", "@r0\nAGCATCGATC\r\n+\nIIIIIIIIII\n@r1\nTCAGTTTTTGA\r\n+\nIIIIIIIIIII\n", "@r0\nTCAGTTTTTGA\n+\nIIIIIIIIIII\n@r1\nAGCATCGATC\r\n+\nIIIIIIIIII", ["-u", "1"], [(0, 8)])
pcase("Fastq paired 4", "q", "@r0\nAGCATCG\n+\nIIIIIII\n", "@r0\nGATCAAAAACTGA\n+\nIIIIIIIIIIIII\n", ["-3", "7"], [])
pcase("Fasta paired 1", "f", "\n\n\r\n>r0\nAGCATCGATC\r\n\n\n>r1\nTCAGTTTTTGA\r\n", "\n\n\r\n>r0\nTCAGTTTTTGA\n\n\n\r\n>r1\nAGCATCGATC", pairhits=[(0, 8), (0, 8)])
pcase("Fasta paired 2", "f", ">r0\nAGCATCGATC\r\n\n\n>r1\nTCAGTTTTTGA\n", "\n\n\r\n>r0\nTCAGTTTTTGA\n\n\n\r\n>r1\nAGCATCGATC", ["-s", "1"], [(0, 8)])
pcase("Fasta paired 3", "f", "\n\n\r\n>r0\nAGCATCGATC\r\n\n\n>r1\nTCAGTTTTTGA\r\n", "\n\n\r\n>r0\nTCAGTTTTTGA\n\n\n\r\n>r1\nAGCATCGATC", ["-u", "1"], [(0, 8)])
pcase("Fasta paired 4", "f", ">\nAGCATCG\n", ">\nGATCAAAAACTGA\n", ["-3", "7"], [])
pcase("Raw paired 1", "r", "\n\n\r\nAGCATCGATC\r\n\n\nTCAGTTTTTGA\r\n", "\n\n\r\nTCAGTTTTTGA\n\n\n\r\nAGCATCGATC", pairhits=[(0, 8), (0, 8)])
pcase("Raw paired 2", "r", "AGCATCGATC\r\n\n\nTCAGTTTTTGA\n", "\n\n\r\nTCAGTTTTTGA\n\n\n\r\nAGCATCGATC", ["-s", "1"], [(0, 8)])
pcase("Raw paired 3", "r", "\n\n\r\nAGCATCGATC\r\n\n\nTCAGTTTTTGA\r\n", "\n\n\r\nTCAGTTTTTGA\n\n\n\r\nAGCATCGATC", ["-u", "1"], [(0, 8)])
pcase("Raw paired 4", "r", "\nAGCATCG\n", "\nGATCAAAAACTGA\n", ["-3", "7"], [])


@pytest.fixture(scope="module")
def env_cpu():
    ensure_oracle_built()
    if not REF_ALIGN.exists() or not REF_BUILD.exists():
        pytest.skip("reference binaries not available")
    import bowtie_b200
    bowtie_b200.build_library()
    from test_cli_parity import build_shim
    build_shim()
    return dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))


_built = {}
def index_for(ref, tmp_root):
    if ref not in _built:
        d = tmp_root / f"ix{len(_built)}"
        d.mkdir()
        (d / "ref.fa").write_text(">0\n" + ref + "\n")
        p = subprocess.run([str(REF_BUILD), "-q", str(d / "ref.fa"), str(d / "ref")], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        _built[ref] = d / "ref"
    return _built[ref]


def run_one(exe, base, kind, text, args, sam, out, env=None, ref=False):
    flags = list(args) + (["-S", "--sam-nohead"] if sam else [])
    if kind == "c":
        src = ["-c", text]
    else:
        f = out.parent / f"reads.{kind}"
        f.write_bytes(text.encode())
        src = [{"q": "-q", "f": "-f", "r": "-r"}[kind], str(f)]
    cmd = [str(exe), *flags] + (["-p", "1"] if ref else []) + ["-x", str(base), *src, str(out)]
    if out.exists():
        out.unlink()
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    body = out.read_bytes() if out.exists() else b""
    summ = "\n".join(l for l in p.stderr.splitlines() if l.startswith("#") or l.startswith("Reported") or l.startswith("No alignments"))
    return p.returncode, body, summ


def check(i, sam, env, tmp_path_factory):
    name, ref, kind, text, args, hits = C[i]
    root = tmp_path_factory.getbasetemp()
    base = index_for(ref, root)
    d = root / f"case{i}_{int(sam)}"
    d.mkdir(exist_ok=True)
    rc_r, body_r, sum_r = run_one(REF_ALIGN, base, kind, text, args, sam, d / "ref.out", ref=True)
    rc_o, body_o, sum_o = run_one(CLI, base, kind, text, args, sam, d / "our.out", env=env)
    if hits is None:
        assert rc_r != 0, "transcription: the reference was expected to abort"
        assert rc_o != 0
        return
    assert rc_r == 0 and rc_o == 0
    if not sam:
        # transcription check against the expected offsets of simple_tests.pl
        got = {}
        for line in body_r.decode().splitlines():
            f = line.split("\t")
            got.setdefault(f[0], {}).setdefault(int(f[3]), 0)
            got[f[0]][int(f[3])] += 1
        want = [h for h in hits if h]
        assert sorted(map(lambda x: sorted(x.items()), got.values())) == sorted(map(lambda x: sorted(x.items()), want)), (name, got, hits)
    assert body_o == body_r, name
    assert sum_o == sum_r, name


@pytest.mark.parametrize("i", range(len(C)), ids=[c[0] for c in C])
@pytest.mark.parametrize("sam", [False, True], ids=["default", "sam"])
def test_simple_case_host_logic(i, sam, env_cpu, tmp_path_factory):
    check(i, sam, env_cpu, tmp_path_factory)


def run_pair(exe, base, kind, t1, t2, args, sam, out, env=None, ref=False):
    flags = list(args) + (["-S", "--sam-nohead"] if sam else [])
    if kind == "c":
        src = ["-c", "-1", t1, "-2", t2]
    else:
        f1, f2 = out.parent / f"m1.{kind}", out.parent / f"m2.{kind}"
        f1.write_bytes(t1.encode()); f2.write_bytes(t2.encode())
        src = [{"q": "-q", "f": "-f", "r": "-r"}[kind], "-1", str(f1), "-2", str(f2)]
    if out.exists():
        out.unlink()
    p = subprocess.run([str(exe), *flags] + (["-p", "1"] if ref else []) + ["-x", str(base), *src, str(out)], capture_output=True, text=True, env=env)
    body = out.read_bytes() if out.exists() else b""
    summ = "\n".join(l for l in p.stderr.splitlines() if l.startswith("#") or l.startswith("Reported") or l.startswith("No alignments"))
    return p.returncode, body, summ


def check_pair(i, sam, env, tmp_path_factory):
    name, ref, kind, t1, t2, args, pairhits = P[i]
    root = tmp_path_factory.getbasetemp()
    base = index_for(ref, root)
    d = root / f"pcase{i}_{int(sam)}"
    d.mkdir(exist_ok=True)
    rc_r, body_r, sum_r = run_pair(REF_ALIGN, base, kind, t1, t2, args, sam, d / "ref.out", ref=True)
    rc_o, body_o, sum_o = run_pair(CLI, base, kind, t1, t2, args, sam, d / "our.out", env=env)
    assert rc_r == 0 and rc_o == 0, (name, rc_r, rc_o)
    if not sam:
        lines = body_r.decode().splitlines()
        got = [tuple(sorted((int(lines[k].split("\t")[3]), int(lines[k + 1].split("\t")[3])))) for k in range(0, len(lines), 2)]
        assert sorted(got) == sorted(pairhits), (name, got, pairhits)      # pins the transcription
    assert body_o == body_r, name
    assert sum_o == sum_r, name


@pytest.mark.parametrize("i", range(len(P)), ids=[c[0] for c in P])
@pytest.mark.parametrize("sam", [False, True], ids=["default", "sam"])
def test_simple_paired_case_host_logic(i, sam, env_cpu, tmp_path_factory):
    check_pair(i, sam, env_cpu, tmp_path_factory)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(P)), ids=[c[0] for c in P])
def test_simple_paired_case_gpu(i, tmp_path_factory):
    ensure_oracle_built()
    if not REF_ALIGN.exists() or not REF_BUILD.exists():
        pytest.skip("reference binaries not available")
    check_pair(i, False, {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"}, tmp_path_factory)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(C)), ids=[c[0] for c in C])
def test_simple_case_gpu(i, env_cpu, tmp_path_factory):
    env = {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"}
    check(i, False, env, tmp_path_factory)
