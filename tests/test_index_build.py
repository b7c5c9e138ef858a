"""Index construction (bt_index_build / bowtie-b200-build): the six index files must be byte-identical to what
the reference's bowtie-build writes for the same FASTA input, -o and -t.

CPU suite: the product's host code (bt_build.h: FASTA records, joined text, side packing, ftab/eftab, file layout) and its
suffix-sort algorithm (bt_build_sa.cuh: prefix doubling) run over the test-only host backend (tests/host_emu/bsa_host.h,
std:: algorithms in place of CUB) through the emulation shim.  GPU suite: the same through libbowtie_b200.so, i.e. with the CUB
backend of bt_build.cu on the device.
"""
import os
import random
import subprocess
from pathlib import Path

import pytest

from helpers import FIXTURES, ROOT, ensure_oracle_built, have_reference
from test_cli_parity import SHIM_DIR, build_shim

REF_BUILD = ROOT / "oracle" / "_ref" / "bowtie-build-s"
CLI = ROOT / "bowtie_b200" / "bowtie-b200-build"
EXTS = ("1", "2", "3", "4", "rev.1", "rev.2")


@pytest.fixture(scope="module")
def tools():
    ensure_oracle_built()
    if not have_reference() or not REF_BUILD.exists():
        pytest.skip("reference bowtie-build not available")
    import bowtie_b200
    bowtie_b200.build_library()
    assert CLI.exists()
    return CLI


def random_genome(rng, td, tag):
    """Several sequences over one or two files: gaps (N, IUPAC codes, '-'), lower case, leading/trailing gaps, repeats that are
    longer than the sort's first key, line widths, empty names, a last line without a newline."""
    nseq = rng.randint(1, 5)
    recs = []
    for k in range(nseq):
        L = rng.choice([5, 30, 200, 1000, 5000])
        unit = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 40)))
        s = []
        while len(s) < L:
            r = rng.random()
            if r < 0.1:
                s += ["N"] * rng.randint(1, 12)
            elif r < 0.15:
                s += [rng.choice("RYKMSWBDHVX-")]
            elif r < 0.3:
                s += list(unit)
            else:
                s += [rng.choice("ACGTacgt") for _ in range(rng.randint(1, 60))]
        if len(s) > 200 and rng.random() < 0.7:
            for _ in range(rng.randint(1, 4)):
                a = rng.randrange(0, len(s) - 100)
                s += s[a:a + rng.randint(30, min(2000, len(s) - a))]
        if rng.random() < 0.2:
            s += list(rng.choice("ACGT") * rng.randint(50, 3000))
        if rng.random() < 0.3:
            s = ["N"] * rng.randint(1, 5) + s
        if rng.random() < 0.3:
            s = s + ["N"] * rng.randint(1, 5)
        if not any(c in "ACGTacgt" for c in s):
            s.append("A")
        seq = "".join(s)
        w = rng.choice([50, 60, 70, 100000])
        name = rng.choice([f"seq{k}", f"seq{k} with words", ""])
        recs.append(f">{name}\n" + "\n".join(seq[i:i + w] for i in range(0, len(seq), w)) + ("\n" if rng.random() < 0.9 else ""))
    nfiles = rng.randint(1, min(2, nseq))
    per = (len(recs) + nfiles - 1) // nfiles
    files = []
    for f in range(nfiles):
        part = recs[f * per:(f + 1) * per]
        if not part:
            continue
        p = td / f"{tag}_{f}.fa"
        p.write_text("".join(x if x.endswith("\n") or i == len(part) - 1 else x + "\n" for i, x in enumerate(part)))
        files.append(str(p))
    return files


def build_both(cli, env, files, td, tag, o, t):
    a = subprocess.run([str(cli), "-q", "-o", str(o), "-t", str(t), ",".join(files), str(td / f"{tag}_ours")], capture_output=True, text=True, env=env)
    b = subprocess.run([str(REF_BUILD), "-q", "-o", str(o), "-t", str(t), ",".join(files), str(td / f"{tag}_ref")], capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    assert a.returncode == 0, a.stderr
    for e in EXTS:
        assert (td / f"{tag}_ours.{e}.ebwt").read_bytes() == (td / f"{tag}_ref.{e}.ebwt").read_bytes(), (tag, e, files, o, t)


def check_builder(cli, env, tmp_path, n_random, genome=True):
    rng = random.Random(20240924)
    for it in range(n_random):
        files = random_genome(rng, tmp_path, f"g{it}")
        build_both(cli, env, files, tmp_path, f"r{it}", rng.choice([1, 3, 5, 7]), rng.choice([2, 5, 7, 10]))
    if genome:
        fna = next((p for p in (FIXTURES / "NC_008253.fna", Path("/root/reference/genomes/NC_008253.fna")) if p.exists()), None)
        if fna is not None:                                   # the shipped e_coli index was built with -t 7
            build_both(cli, env, [str(fna)], tmp_path, "ecoli", 5, 7)
            for e in ("1", "2", "rev.1", "rev.2"):
                assert (tmp_path / f"ecoli_ours.{e}.ebwt").read_bytes() == (FIXTURES / f"e_coli.{e}.ebwt").read_bytes()


def test_index_files_match_bowtie_build_host_emulation(tools, tmp_path):
    build_shim()
    check_builder(tools, dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR)), tmp_path, 25)


def test_build_cli_command_line_sequences_and_errors(tools, tmp_path):
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    seqs = "ACGTTGCANNACGTAGCTAGCTAGGATCGAT,GGGATTTAGGCATACGATCCCAGATAGGACCATTTAGAGAGCCCAT"
    a = subprocess.run([str(tools), "-q", "-c", "-t", "4", seqs, str(tmp_path / "c_ours")], capture_output=True, text=True, env=env)
    b = subprocess.run([str(REF_BUILD), "-q", "-c", "-t", "4", seqs, str(tmp_path / "c_ref")], capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    for e in EXTS:
        assert (tmp_path / f"c_ours.{e}.ebwt").read_bytes() == (tmp_path / f"c_ref.{e}.ebwt").read_bytes(), e
    bad = tmp_path / "bad.fa"
    bad.write_text("ACGT\n")
    p = subprocess.run([str(tools), "-q", str(bad), str(tmp_path / "x")], capture_output=True, text=True, env=env)
    assert p.returncode != 0 and "FASTA" in p.stderr
    p = subprocess.run([str(tools), "-q", "--ntoa", str(bad), str(tmp_path / "x")], capture_output=True, text=True, env=env)
    assert p.returncode != 0 and "ntoa" in p.stderr
    p = subprocess.run([str(tools), "-q", str(tmp_path / "missing.fa"), str(tmp_path / "x")], capture_output=True, text=True, env=env)
    assert p.returncode != 0


def test_built_index_is_searchable(tools, tmp_path):
    """An index written by the builder goes through the loader and the search like one written by bowtie-build."""
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))
    rng = random.Random(7)
    files = random_genome(rng, tmp_path, "s")
    build_both(tools, env, files, tmp_path, "s", 3, 6)
    align = ROOT / "bowtie_b200" / "bowtie-b200-align"
    text = "".join(l.strip() for f in files for l in Path(f).read_text().splitlines() if not l.startswith(">"))
    clean = [text[i:i + 30] for i in range(0, len(text) - 30, 37) if all(c in "ACGTacgt" for c in text[i:i + 30])][:20]
    if not clean:
        pytest.skip("no gap-free windows in this genome")
    outs = []
    for base in ("s_ours", "s_ref"):
        p = subprocess.run([str(align), "-v", "1", "-a", "-x", str(tmp_path / base), "-c", ",".join(clean), str(tmp_path / f"{base}.hits")], capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        outs.append((tmp_path / f"{base}.hits").read_bytes())
    assert outs[0] == outs[1] and len(outs[0]) > 0


def check_build_text(lib_env, tmp_path, total_len):
    """bt_index_build_text on bench.py's synthetic genome (records with N gaps, no FASTA) against bowtie-build on the FASTA
    written from the same records: the entry point bench.py builds its hg19-sized index with."""
    p = subprocess.run([os.sys.executable, str(ROOT / "tests" / "build_text_child.py"), str(total_len), str(tmp_path / "bt_ours"), str(tmp_path / "bt.fa")],
                       capture_output=True, text=True, env=lib_env)
    assert p.returncode == 0, p.stderr
    b = subprocess.run([str(REF_BUILD), "-q", "-o", "4", "-t", "8", str(tmp_path / "bt.fa"), str(tmp_path / "bt_ref")], capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    for e in EXTS:
        assert (tmp_path / f"bt_ours.{e}.ebwt").read_bytes() == (tmp_path / f"bt_ref.{e}.ebwt").read_bytes(), e


def test_build_text_matches_bowtie_build_host_emulation(tools, tmp_path):
    build_shim()
    check_build_text(dict(os.environ, BOWTIE_B200_LIB=str(SHIM_DIR / "libbowtie_b200.so")), tmp_path, 600_000)


@pytest.mark.gpu
def test_index_files_match_bowtie_build_gpu(tools, tmp_path):
    """The CUB backend on the device: e_coli against the reference's shipped index files, 25 random multi-sequence genomes and
    the in-memory entry point against bowtie-build."""
    env = {k: v for k, v in os.environ.items() if k not in ("LD_LIBRARY_PATH", "BOWTIE_B200_LIB")}
    check_builder(tools, env, tmp_path, 25)
    check_build_text(env, tmp_path, 3_000_000)
