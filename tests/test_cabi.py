"""The C-ABI shared library loads and exports exactly the symbols include/kdsl.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "kdsl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kdsl_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from kagomedsl.jl_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"libkdsl.so does not export {s}"
    assert sorted(_lib.SYMBOLS) == syms, "python binding list out of sync with include/kdsl.h"
    assert L.kdsl_version() >= 100


def test_header_is_plain_c():
    txt = open(os.path.join(ROOT, "include", "kdsl.h")).read()
    assert 'extern "C"' in txt and "torch" not in txt.lower() and "std::" not in txt
    # compiles as C
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "kdsl.h"\nint main(void){ kdsl_handle h = 0; (void)h; return KDSL_N_ACC == 8 ? 0 : 1; }\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", os.path.join(d, "t.o")])


def test_error_reporting_without_gpu():
    from kagomedsl.jl_b200 import _lib
    L = _lib.lib()
    n = ctypes.c_int(-1)
    rc = L.kdsl_device_count(ctypes.byref(n))
    if rc == 0:
        assert n.value >= 1
    else:
        assert rc == _lib.KDSL_ERR_CUDA and n.value == 0 and len(L.kdsl_last_error()) > 0
    assert L.kdsl_destroy(None) == 0
    assert L.kdsl_set_sweeps(None, 0) == _lib.KDSL_ERR_INVALID_ARGUMENT


def test_sass_has_dmma_and_128bit_streaming():
    """evidence the shipped binary is sm_100a code using the FP64 tensor pipe and 128-bit global access"""
    import shutil, subprocess
    from kagomedsl.jl_b200 import _lib
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out.upper() or "sm_100" in out
    assert "DMMA" in out
    assert re.search(r"LDG\.E(\.NA)?\.128", out) and re.search(r"STG\.E(\.NA)?\.128", out)
