"""The best-first ("stateful") search path — --best, --strata, -M, -v 3 — against the
unmodified reference binary (oracle/_ref/bowtie-align-s), byte for byte, on synthetic reads that stress what the
shipped e_coli reads do not: Ns, qualities below the Maq rounding threshold (zero-cost mismatches), ragged
lengths from 4 bases up, repeats, and many seedlings per read.

* not-gpu: through tests/host_emu/shim (the device code compiled for the host) — checks the engine's logic.
* gpu: the CUDA library, including the passes that re-run reads whose 64 KB arena overflowed.
"""
import os
import subprocess
from pathlib import Path

import pytest

from helpers import REF_ALIGN, ROOT, ensure_oracle_built, have_reference
from test_cli_parity import CLI, SHIM_DIR, build_shim

FLAGS = [
    ["-n", "2", "--best"],
    ["-n", "3", "--best", "--strata", "-a"],
    ["-n", "1", "--best", "-k", "4"],
    ["-v", "3", "-k", "3"],
    ["-v", "2", "--best", "--strata", "-k", "2"],
    ["-n", "2", "--best", "--nomaqround", "-e", "200", "-k", "5"],
    ["-n", "2", "-M", "2"],
    ["-n", "2", "--best", "-l", "12", "-k", "3", "--strata"],
    ["-n", "2", "--best", "-e", "300", "-k", "6"],
]


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    ensure_oracle_built()
    if not have_reference():
        pytest.skip("reference binary / fixtures not available")
    import bowtie_b200
    bowtie_b200.build_library()
    from synth import build_synth_index, synth_reads, write_fastq
    base, genome = build_synth_index("t1", 3, 300000, 7)
    d = tmp_path_factory.mktemp("bf")
    write_fastq(d / "low.fq", synth_reads(genome, 500, (18, 120), seed=3, sub_rate=0.03, n_rate=0.005, qual_profile="low"))
    write_fastq(d / "short.fq", synth_reads(genome, 300, (4, 40), seed=8, sub_rate=0.05, n_rate=0.02, qual_profile="low"))
    write_fastq(d / "big.fq", synth_reads(genome, 20000, 100, seed=5, sub_rate=0.02, qual_profile="mixed"))
    return base, d


def run(exe, flags, base, reads, out, env=None, ref_threads=0):
    cmd = [str(exe), *flags] + (["-p", "1"] if ref_threads else []) + ["-x", str(base), str(reads), str(out)]   # one thread: output in read order
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert p.returncode == 0, p.stderr
    summary = "\n".join(l for l in p.stderr.splitlines() if l.startswith("#") or l.startswith("Reported") or l.startswith("No alignments"))
    assert "Exhausted best-first chunk memory" not in p.stderr     # the reference's own memory limit: not a parity case
    return Path(out).read_bytes(), summary


def compare(flags, base, reads, tmp_path, env, ref_threads=1):
    ref = run(REF_ALIGN, flags, base, reads, tmp_path / "ref.out", ref_threads=ref_threads)
    ours = run(CLI, flags, base, reads, tmp_path / "our.out", env=env)
    assert ours[0] == ref[0]
    assert ours[1] == ref[1]
    assert len(ref[0]) > 0


@pytest.mark.parametrize("flags", FLAGS, ids=["_".join(f).replace("--", "") for f in FLAGS])
@pytest.mark.parametrize("reads", ["low.fq", "short.fq"])
def test_best_first_logic_matches_reference(flags, reads, setup, tmp_path):
    base, d = setup
    build_shim()
    compare(flags, base, d / reads, tmp_path, dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR)))


def test_best_first_small_arena_is_rerun(setup, tmp_path):
    """Reads that exhaust their arena are searched again with a larger one; the result must not depend on the arena size."""
    base, d = setup
    build_shim()
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR), BT_EMU_ARENA_WORDS="3000")
    compare(["-n", "2", "--best", "-k", "3"], base, d / "low.fq", tmp_path, env)
    # and nothing may depend on the arena being zeroed (the GPU arenas never are): fill it with a pattern before every read
    env = dict(os.environ, LD_LIBRARY_PATH=str(SHIM_DIR), BT_EMU_ARENA_POISON="600000")
    compare(["-n", "3", "--best", "--strata", "-a"], base, d / "low.fq", tmp_path, env)


def gpu_env(**kw):
    env = {k: v for k, v in os.environ.items() if k != "LD_LIBRARY_PATH"}
    env.update(kw)
    return env


@pytest.mark.gpu
@pytest.mark.parametrize("flags", FLAGS, ids=["_".join(f).replace("--", "") for f in FLAGS])
@pytest.mark.parametrize("reads", ["low.fq", "short.fq"])
def test_best_first_gpu_matches_reference(flags, reads, setup, tmp_path):
    base, d = setup
    compare(flags, base, d / reads, tmp_path, gpu_env())


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [["-n", "2", "--best"], ["-v", "3", "-k", "2"], ["-n", "2", "--best", "--strata", "-k", "3"]], ids=["n2best", "v3k2", "n2strata"])
def test_best_first_gpu_20k_reads(flags, setup, tmp_path):
    base, d = setup
    compare(flags, base, d / "big.fq", tmp_path, gpu_env(), ref_threads=1)


@pytest.mark.gpu
def test_best_first_gpu_arena_tiers(setup, tmp_path):
    """A 4 KB first-tier arena sends most reads through the 1 MB (and some through the 16 MB) pass: same bytes out."""
    base, d = setup
    compare(["-n", "2", "--best", "-k", "3"], base, d / "low.fq", tmp_path, gpu_env(BT_BEST_ARENA_KW="1"))
    compare(["-n", "3", "--best", "-a", "--strata"], base, d / "big.fq", tmp_path, gpu_env(BT_BEST_ARENA_KW="2"), ref_threads=1)


def test_fuzz_both_paths_against_reference(setup):
    """tools/fuzz_cli.py (random genomes / reads / option sets, both search paths) — a short run as a regression test."""
    build_shim()
    p = subprocess.run([str(ROOT / "tools" / "fuzz_cli.py"), "--iters", "25", "--seed", "11"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-3000:]


@pytest.mark.gpu
def test_fuzz_both_paths_gpu(setup):
    p = subprocess.run([str(ROOT / "tools" / "fuzz_cli.py"), "--iters", "12", "--seed", "12", "--gpu"], capture_output=True, text=True, env=gpu_env())
    assert p.returncode == 0, p.stdout[-3000:]
