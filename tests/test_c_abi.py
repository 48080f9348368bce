"""The C-ABI library loads and exports every symbol include/scbonita_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from common import CUDA_LIB, ROOT
from scbonita_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "scbonita_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*\s*|int\s+|void\s+)(\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(CUDA_LIB), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(CUDA_LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", CUDA_LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_device_means_loud_failure():
    """Without a GPU every compute entry point must fail with a CUDA error, never fall back to the CPU."""
    lib = _lib.load()
    if lib.scb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import numpy as np
    ctx = ctypes.c_void_p()
    bm = np.zeros((2, 32), np.int32)
    pos = np.arange(2, dtype=np.int32)
    rc = lib.scb_ctx_create(ctypes.byref(ctx), 0, 2, 32, bm.ctypes.data, 4, 32, pos.ctypes.data)
    assert rc == _lib.ERR_CUDA
    assert b"CUDA" in lib.scb_last_error() or b"device" in lib.scb_last_error()


def test_product_never_imports_oracle():
    """The product package must not import, link or execute anything under oracle/ or tests/emu."""
    pkg = os.path.join(ROOT, "scbonita_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
                assert "scb_oracle" not in src, f
                assert "libscbonita_emu" not in src, f
