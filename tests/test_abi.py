"""CPU-side checks of the C ABI: the library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

import mmmfe_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmmfe_[a-z0-9_]+)\s*\(", txt)))


def test_engine_exports_every_declared_symbol():
    pkg = mmmfe_import.load()
    lib = pkg.load_engine()
    names = _declared("mmmfe_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert set(lib._signatures) == set(names)
    assert b"sm_100a" in lib.mmmfe_version()


def test_engine_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pkg = mmmfe_import.load()
    with pytest.raises(pkg.MmmfeError, match="no CPU path"):
        pkg.Engine()


def test_engine_binary_has_dmma_and_no_fallback_arch():
    """sm_100a SASS only, with the FP64 tensor-core instruction present."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    pkg = mmmfe_import.load()
    out = subprocess.run(["cuobjdump", "-lelf", pkg.ENGINE_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.run(["cuobjdump", "-sass", pkg.ENGINE_PATH], capture_output=True, text=True).stdout
    gemm = sass[sass.index("k_grouped_gemm"):]
    gemm = gemm[:gemm.index("Function :", 10)] if "Function :" in gemm[10:] else gemm
    assert "DMMA" in gemm
