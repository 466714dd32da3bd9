"""The C-ABI library loads on a CPU-only box, exports every symbol include/udales_gpu.h declares,
and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import udales_b200 as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "udales_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(udgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    if not os.path.exists(U.LIB_PATH):
        U.build()
    L = ctypes.CDLL(U.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(U.EXPORTS) <= set(names)


def test_abi_version_and_struct_size():
    L = U.lib()
    assert L.udgpu_abi_version() == U.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(U.UdalesGPUError) as e:
        U.UdalesGPU(8, 8, 8)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/."""
    pkg = os.path.join(ROOT, "u-dales_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|udales_oracle|oracle/|orc_[a-z]+\s*\(")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".f90", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not pat.search(txt), (dp, f, pat.search(txt).group(0))
    txt = open(os.path.join(ROOT, "include", "udales_gpu.h")).read()
    assert not pat.search(txt)
