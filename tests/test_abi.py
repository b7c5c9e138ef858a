"""The C-ABI library builds, loads, and exports every symbol include/bowtie_b200.h declares.
No compute calls here (no GPU on the CPU test box)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import bowtie_b200
    bowtie_b200.build_library()
    return bowtie_b200.load_library()


def declared_functions():
    src = (ROOT / "include" / "bowtie_b200.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bt_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    fns = declared_functions()
    for need in ("bt_index_load", "bt_index_free", "bt_align_batch", "bt_align_batch_device", "bt_policy_init", "bt_last_error"):
        assert need in fns


def test_library_exports_every_declared_symbol(lib):
    for fn in declared_functions():
        assert hasattr(lib, fn), fn


def test_abi_version_and_policy_defaults(lib):
    from bowtie_b200.api import _Policy
    assert lib.bt_abi_version() == 5
    p = _Policy()
    lib.bt_policy_init(C.byref(p))
    # resetOptions defaults (ebwt_search.cpp:181-219)
    assert (p.mode, p.mms, p.seed_len, p.qual_thresh, p.max_bts, p.khits, p.mhits, p.maq_round) == (1, 2, 28, 70, 125, 1, 0xFFFFFFFF, 1)


def test_no_cpu_search_path(lib, tmp_path):
    """Without a CUDA device the product must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.bt_index_load(str(tmp_path / "nope").encode(), 0, 0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"CUDA" in lib.bt_last_error() or b"cuda" in lib.bt_last_error()


def test_product_does_not_reference_the_oracle():
    for p in (ROOT / "bowtie_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and p.is_file():
            txt = p.read_text(errors="ignore")
            assert "bt_oracle" not in txt and "libbtoracle" not in txt and "host_emu/" not in txt.replace("tests/host_emu/", ""), p
