"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
declared in include/hsk_capi.h (no compute calls without a GPU)."""
import ctypes
import os
import re

from hysortk_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hsk_capi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hsk_[a-z_]+)\s*\(", text)))


def test_library_builds_and_exports_all_symbols():
    so = build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    syms = declared_symbols()
    assert set(syms) == set(capi.EXPORTS), (syms, capi.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in hsk_capi.h but not exported"
    assert lib.hsk_version() == 2


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.HskError):
        capi.Context(31, 17, 2, 50)


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", capi.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out
