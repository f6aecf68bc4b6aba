"""The C-ABI library: builds on the CPU box, loads, exports every symbol the headers declare, and fails LOUDLY
(no CPU fallback) when asked to compute without a GPU.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"SVDLAS2_API[^;(]*?\b(svdlas2_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c():
    """include/svdlas2.h compiles as C11 with no other include path (plain pointers and sizes, no torch types)"""
    src = '#include "svdlas2.h"\n#include "svdlas2_testing.h"\nint main(void){return svdlas2_abi_version()==SVDLAS2_ABI_VERSION?0:1;}\n'
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-"], input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_every_declared_symbol_is_exported(las2):
    from svdlibrs_b200 import _lib
    names = declared("svdlas2.h")
    assert names == sorted(_lib.SYMBOLS), "svdlibrs_b200/_lib.py SYMBOLS is out of sync with include/svdlas2.h"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (svdlas2_[a-z0-9_]+)", out))
    for n in names + declared("svdlas2_testing.h"):
        assert n in exported, f"{n} is declared in include/ but not exported by the library"
    L = _lib.lib()
    assert L.svdlas2_abi_version() == 2
    assert L.svdlas2_last_error() == b""


def test_library_is_sm100a_and_has_no_cpu_path(las2):
    from svdlibrs_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    # nothing under oracle/ is linked or referenced by the product
    deps = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    for root, _, files in os.walk(os.path.join(ROOT, "svdlibrs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                assert "oracle" not in open(os.path.join(root, f)).read().lower().replace("oracle/", "ORACLE_DIR") \
                    or f in ("__init__.py",), f


def test_struct_layouts_match(las2):
    """ctypes mirrors vs the C compiler's view of the header"""
    from svdlibrs_b200 import _lib
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "svdlas2.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
           'sizeof(svdlas2_diagnostics),sizeof(svdlas2_profile),sizeof(svdlas2_result),'
           'offsetof(svdlas2_result,profile),offsetof(svdlas2_result,trace_len));return 0;}\n')
    exe = os.path.join(ROOT, "tests", ".layout_probe")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src.encode(), check=True)
    try:
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    finally:
        os.remove(exe)
    assert got == [C.sizeof(_lib.Diagnostics), C.sizeof(_lib.Profile), C.sizeof(_lib.Result),
                   _lib.Result.profile.offset, _lib.Result.trace_len.offset]


def test_compute_without_gpu_fails_loudly(las2):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(las2.SvdLibError) as e:
        las2.svd(sp.csr_matrix(np.eye(3)))
    assert e.value.kind == "CudaError" and "no CPU fallback" in str(e.value)
    with pytest.raises(las2.SvdLibError):
        las2.Operator(sp.csr_matrix(np.eye(3)))


def test_argument_validation_needs_no_gpu(las2):
    from svdlibrs_b200 import _lib
    L = _lib.lib()
    assert L.svdlas2_operator_solve(None, 2, 0, None, 1e-6, 1, None) == 6
    assert b"null" in L.svdlas2_last_error()
    with pytest.raises(TypeError):
        las2.svd(np.eye(3))  # dense input is not an SMat
    with pytest.raises(las2.SvdLibError) as e:  # 1 x k: rejected before any device work (src/lib.rs:596-600)
        las2.svd(sp.csr_matrix(np.ones((1, 5))))
    assert e.value.kind == "Las2Error" and str(e.value) == "svdlibrs/svdLas2: svdLAS2: insufficient dimensions: 1"
