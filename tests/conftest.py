import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from helpers import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ecoli_base():
    from helpers import FIXTURES, have_reference, ensure_oracle_built
    ensure_oracle_built()
    if not have_reference():
        pytest.skip("oracle/_ref fixtures not available (built from /root/reference by oracle/Makefile)")
    return FIXTURES / "e_coli"


@pytest.fixture(scope="session")
def ecoli_reads():
    from helpers import FIXTURES, parse_fastq, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref fixtures not available")
    return parse_fastq(FIXTURES / "e_coli_1000.fq")


@pytest.fixture(scope="session")
def synth_index():
    """Small multi-sequence index with N gaps, built by the reference's own bowtie-build."""
    from synth import build_synth_index
    from helpers import REF_BUILD
    if not REF_BUILD.exists():
        pytest.skip("oracle/_ref/bowtie-build-s not available")
    return build_synth_index("t1", n_seqs=3, total_len=300_000, seed=7, ftab_chars=10, with_gaps=True)
