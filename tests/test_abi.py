"""The C-ABI library loads and exports every symbol include/tdm_b200.h declares (no compute)."""
import ctypes
import subprocess
import sys
from pathlib import Path

import pytest

from tinydiffusionmodels_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not _lib.LIB_PATH.exists():
        subprocess.check_call([sys.executable, "-m", "tinydiffusionmodels_b200.build"], cwd=ROOT)
    return _lib.load()


def test_header_and_bindings_agree(lib):
    declared = set(_lib.header_symbols())
    bound = set(_lib.SIGNATURES)
    assert declared == bound, f"header-only: {declared - bound}; binding-only: {bound - declared}"


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in _lib.header_symbols():
        assert hasattr(raw, name), f"{name} is declared in include/tdm_b200.h but not exported"


def test_metadata_calls(lib):
    assert lib.tdm_version() == 1
    assert lib.tdm_unet_param_count() == 181_473
    assert lib.tdm_unet_wpack_bytes() > 181_473 * 4
    assert lib.tdm_unet_workspace_bytes(64, 0) > 0
    assert lib.tdm_last_error() is not None


def test_argument_errors_are_reported_without_a_gpu(lib):
    # argument validation happens before any CUDA call: inner % 4 != 0 is rejected
    rc = lib.tdm_q_sample(16, 16, 16, 16, 16, 16, 1, 3, 1000, None)
    assert rc == 1
    assert b"multiple of 4" in lib.tdm_last_error()


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package or src/ may import it."""
    for base in ("tinydiffusionmodels_b200", "src"):
        for f in (ROOT / base).rglob("*.py"):
            text = f.read_text()
            assert "import oracle" not in text and "from oracle" not in text, f
