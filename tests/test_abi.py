"""The C-ABI library loads and exports every symbol include/divergen_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "divergen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    from divergen_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and header disagree"
    assert built_lib.dg_version() >= 100


def test_cpu_only_fails_loudly(built_lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = built_lib.dg_ctx_create(0, C.byref(h))
    assert rc != 0 and built_lib.dg_last_error()
    from divergen_b200 import UNet2DConditionModel
    with pytest.raises((RuntimeError, ValueError)):
        UNet2DConditionModel(device="cuda:0")
    with pytest.raises(ValueError):
        UNet2DConditionModel(device="cpu")


def test_sass_is_blackwell_native():
    """cuobjdump evidence: tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld/st -> LDTM/STTM; no legacy HMMA."""
    import shutil
    import subprocess
    from divergen_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or library unavailable")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for m in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert m in sass, m
    assert not re.search(r"\bHMMA\b", sass)
