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


def test_tensor_core_kernels_carry_tcgen05_and_bulk_copy_sass():
    """Static proof that the hot kernels are tcgen05 / TMEM / TMA-engine code and not a SIMT or mma.sync stand-in:
    every convolution, weight-gradient, GEMM, FFN and attention kernel in the built library issues UTCHMMA
    (tcgen05.mma), reads its accumulator with LDTM (tcgen05.ld) and is fed by UBLKCP (cp.async.bulk)."""
    import shutil
    import sys
    from pathlib import Path

    if not (shutil.which("cuobjdump") and shutil.which("c++filt")):
        pytest.skip("cuobjdump / c++filt not on PATH")
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    from sass_mnemonics import kernel_mnemonics

    table = kernel_mnemonics()
    hot = {n: c for n, c in table.items()
           if n.split("<")[0] in ("conv3x3_tc_kernel", "wgrad_dup_kernel", "wgrad_tc_kernel", "gemm_tc_kernel",
                                  "ffn_tc_kernel", "attn_tc_kernel")}
    assert len(hot) >= 30
    for name, c in hot.items():
        assert c["UTCHMMA"] > 0 and c["LDTM"] > 0 and c["UBLKCP"] > 0 and c["UTCBAR"] > 0, (name, dict(c))
    assert not any("HMMA." in n for n in table)      # sanity: names are kernels, not instructions
