"""Paired-end alignment (BASELINE config 5): PairedBWAlignerV1 over the best-first drivers, the
reference-window scan for the opposite mate (RefAligner family) and the bit-pair reference (X.3/X.4.ebwt), through the
bowtie-compatible driver, byte for byte against the unmodified reference binary.

* not-gpu: through tests/host_emu/shim (device code compiled for the host).
* gpu: the CUDA library.
"""
import hashlib
import os
import subprocess
from pathlib import Path

import pytest

from helpers import FIXTURES, REF_ALIGN, ROOT, ensure_oracle_built, have_reference
from test_cli_parity import CLI, SHIM_DIR, build_shim

# md5 of the reference's hit file for the shipped paired reads
GOLDEN = {"-n 3": "13330dbc5beb1b9b3070d6c6939ac9b7", "-n 2": "13330dbc5beb1b9b3070d6c6939ac9b7", "-v 2": "956cd5667fdec6116b3672ea1eac78b2",
          "-n 3 --best": "95df699b45a45d4a4c6072e7b0361463"}

ECOLI_FLAGS = [
    "-n 3", "-n 2", "-n 1", "-n 0", "-v 0", "-v 1", "-v 2", "-v 3", "-n 2 -k 3", "-n 2 -a", "-n 2 -m 1", "-n 2 -m 2 -k 2", "-v 2 -a",
    "-n 2 -S", "-v 1 -S -k 2", "-n 2 -X 150", "-n 2 -I 100 -X 300", "-n 2 --ff", "-n 2 --rf", "-n 2 --nofw", "-n 2 --norc",
    "-n 2 -l 20 -e 100", "-n 3 --nomaqround -a", "-n 2 --pairtries 2", "-n 2 -5 2 -3 3", "-n 2 --maxbts 3", "-n 2 -S --no-unal -X 120",
    # PairedBWAlignerV2 (--best) and -M on pairs
    "-n 3 --best", "-n 2 --best", "-v 2 --best", "-n 2 --best --strata -k 3", "-n 2 --best -a", "-v 3 --best -k 2", "-n 2 -M 1", "-n 2 -M 2 --best",
    "-n 2 --best -S", "-n 2 -M 1 -S", "-n 2 --best --strata -a -m 3",
]
SYNTH_FLAGS = ["-n 2", "-n 3 -a", "-n 1 -k 3", "-v 0", "-v 2 -k 2", "-v 3", "-n 2 -m 2", "-n 2 -I 150 -X 260", "-n 2 -l 15 -e 200 -a", "-n 2 -S", "-v 1 -a -X 500",
               "-n 2 --best", "-n 3 --best --strata -a", "-v 3 --best -k 3", "-n 2 -M 2", "-v 1 --best -a -X 400", "-n 1 --best --strata -m 3 -k 2"]


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    ensure_oracle_built()
    if not have_reference():
        pytest.skip("reference binary / fixtures not available")
    import bowtie_b200
    bowtie_b200.build_library()
    from synth import build_synth_index, synth_pairs, write_fastq_pairs
    d = tmp_path_factory.mktemp("pe")
    base, genome = build_synth_index("t1", 3, 300000, 7)
    write_fastq_pairs(d / "low_1.fq", d / "low_2.fq", synth_pairs(genome, 400, (20, 110), seed=21, sub_rate=0.03, n_rate=0.004, qual_profile="low"))
    gbase, ggenome = build_synth_index("t1", 3, 300000, 7, with_gaps=True)
    write_fastq_pairs(d / "gap_1.fq", d / "gap_2.fq", synth_pairs(ggenome, 400, (30, 60), seed=23, sub_rate=0.02, n_rate=0.002, frag_mean=120, frag_sd=30))
    write_fastq_pairs(d / "big_1.fq", d / "big_2.fq", synth_pairs(genome, 10000, 100, seed=25, sub_rate=0.015))
    return {"low": (base, d / "low_1.fq", d / "low_2.fq"), "gap": (gbase, d / "gap_1.fq", d / "gap_2.fq"), "big": (base, d / "big_1.fq", d / "big_2.fq"),
            "ecoli": (FIXTURES / "e_coli", FIXTURES / "e_coli_1000_1.fq", FIXTURES / "e_coli_1000_2.fq")}


def run(exe, flags, case, out, env=None, ref=False):
    base, m1, m2 = case
    cmd = [str(exe), *flags.split()] + (["-p", "1"] if ref else []) + ["-x", str(base), "-1", str(m1), "-2", str(m2), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    body = b"".join(l for l in Path(out).read_bytes().splitlines(keepends=True) if not l.startswith(b"@PG"))
    summary = "\n".join(l for l in p.stderr.splitlines() if l.startswith("#") or l.startswith("Reported") or l.startswith("No alignments"))
    return body, summary


def compare(flags, case, tmp_path, env):
    ref = run(REF_ALIGN, flags, case, tmp_path / "ref.out", ref=True)
    ours = run(CLI, flags, case, tmp_path / "our.out", env=env)
    assert ours[0] == ref[0]
    assert ours[1] == ref[1]
    return ours[0]


def shim_env():
    build_shim()
    return dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR))


def gpu_env():
    return {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"}


@pytest.mark.parametrize("flags", ECOLI_FLAGS, ids=[f.replace(" ", "_") for f in ECOLI_FLAGS])
def test_paired_ecoli_logic(flags, setup, tmp_path):
    body = compare(flags, setup["ecoli"], tmp_path, shim_env())
    if flags in GOLDEN:
        assert hashlib.md5(body).hexdigest() == GOLDEN[flags]


@pytest.mark.parametrize("flags", SYNTH_FLAGS, ids=[f.replace(" ", "_") for f in SYNTH_FLAGS])
@pytest.mark.parametrize("reads", ["low", "gap"])
def test_paired_synthetic_logic(flags, reads, setup, tmp_path):
    compare(flags, setup[reads], tmp_path, shim_env())


@pytest.mark.gpu
@pytest.mark.parametrize("flags", ECOLI_FLAGS, ids=[f.replace(" ", "_") for f in ECOLI_FLAGS])
def test_paired_ecoli_gpu(flags, setup, tmp_path):
    body = compare(flags, setup["ecoli"], tmp_path, gpu_env())
    if flags in GOLDEN:
        assert hashlib.md5(body).hexdigest() == GOLDEN[flags]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", SYNTH_FLAGS, ids=[f.replace(" ", "_") for f in SYNTH_FLAGS])
@pytest.mark.parametrize("reads", ["low", "gap"])
def test_paired_synthetic_gpu(flags, reads, setup, tmp_path):
    compare(flags, setup[reads], tmp_path, gpu_env())


@pytest.mark.gpu
@pytest.mark.parametrize("flags", ["-n 3", "-v 2 -k 2"])
def test_paired_gpu_10k_pairs(flags, setup, tmp_path):
    compare(flags, setup["big"], tmp_path, gpu_env())


def test_interleaved_input(setup, tmp_path):
    """--interleaved: one FASTQ file with alternating mates (pat.cpp) must give what -1/-2 give, and what the reference gives."""
    base, m1, m2 = setup["ecoli"]
    inter = tmp_path / "inter.fq"
    a, b = Path(m1).read_text().splitlines(), Path(m2).read_text().splitlines()
    inter.write_text("".join("\n".join(a[4 * i:4 * i + 4]) + "\n" + "\n".join(b[4 * i:4 * i + 4]) + "\n" for i in range(len(a) // 4)))
    for flags in (["-n", "2"], ["-v", "2", "-k", "2", "-S"], ["-n", "2", "-s", "10", "-u", "300"]):
        outs = []
        for exe, env, extra in ((REF_ALIGN, None, ["-p", "1"]), (CLI, shim_env(), [])):
            out = tmp_path / f"{Path(exe).name}.out"
            p = subprocess.run([str(exe), *flags, *extra, "-x", str(base), "--interleaved", str(inter), str(out)], capture_output=True, text=True, env=env)
            assert p.returncode == 0, p.stderr
            outs.append(b"".join(l for l in out.read_bytes().splitlines(keepends=True) if not l.startswith(b"@PG")))
        assert outs[0] == outs[1] and len(outs[0]) > 0


def make_tabbed(setup, tmp_path, mixed):
    import random
    base, m1, m2 = setup["ecoli"]
    a, b = Path(m1).read_text().splitlines(), Path(m2).read_text().splitlines()
    u = (FIXTURES / "e_coli_1000.fq").read_text().splitlines()
    rng = random.Random(3)
    out = []
    for i in range(600):
        if not mixed or rng.random() < 0.5:
            out.append("\t".join([a[4 * i][1:].replace("/1", ""), a[4 * i + 1], a[4 * i + 3], b[4 * i + 1], b[4 * i + 3]]))
        else:
            out.append("\t".join([u[4 * i][1:], u[4 * i + 1], u[4 * i + 3]]))
        if rng.random() < 0.1:
            out.append("")
    f = tmp_path / "reads.tab"
    f.write_text("\n".join(out) + "\n")
    return base, f


def check_tabbed(setup, tmp_path, env):
    """--12 (TabbedPatternSource): records may mix pairs and single reads; any paired input makes the whole run stateful, and the
    summary reports "N paired-end alignments and M singleton alignments" (hit.h:322-337)."""
    for mixed in (True, False):
        base, f = make_tabbed(setup, tmp_path, mixed)
        for flags in (["-n", "2"], ["-v", "2", "-k", "2", "-S"], ["-n", "2", "-s", "10", "-u", "300", "-5", "2", "-3", "1"], ["-v", "3", "-a"],
                      ["-n", "2", "--best", "--strata", "-k", "3"], ["-n", "2", "-M", "1"]):
            res = []
            for exe, e, extra in ((REF_ALIGN, None, ["-p", "1"]), (CLI, env, [])):
                out = tmp_path / f"{Path(exe).name}.out"
                p = subprocess.run([str(exe), *flags, *extra, "-x", str(base), "--12", str(f), str(out)], capture_output=True, text=True, env=e)
                assert p.returncode == 0, p.stderr
                body = b"".join(l for l in out.read_bytes().splitlines(keepends=True) if not l.startswith(b"@PG"))
                res.append((body, [l for l in p.stderr.splitlines() if l.startswith("#") or l.startswith("Reported")]))
            assert res[0] == res[1], (mixed, flags)
            assert len(res[0][0]) > 0


def test_tabbed_input(setup, tmp_path):
    check_tabbed(setup, tmp_path, shim_env())


@pytest.mark.gpu
def test_tabbed_input_gpu(setup, tmp_path):
    check_tabbed(setup, tmp_path, gpu_env())
