"""The C-ABI library loads without a GPU and exports every symbol include/nav24_orb.h declares;
argument validation that needs no device; the product has no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nav24_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    syms = capi.declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert L.nav24_abi_version() == 2


def test_header_is_plain_c():
    txt = open(capi.HEADER_PATH).read()
    assert 'extern "C"' in txt
    code = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)      # declarations only: comments may cite std::sort etc.
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "cv::" not in code
    # compiles as C
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "nav24_orb.h"\nint main(void){nav24_kp k; return sizeof(k)==28?0:1;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o",
                               os.path.join(d, "t")])
        subprocess.check_call([os.path.join(d, "t")])


def test_struct_layouts_match_header():
    assert capi.KP_DTYPE.itemsize == 28 and C.sizeof(capi.Params) == 24 and C.sizeof(capi.GridCfg) == 24


def test_bad_arguments_without_device():
    L = capi.lib()
    h = C.c_void_p()
    assert L.nav24_orb_create(None, 0, C.byref(h)) == capi.E_BADARG
    prm = capi.Params(1000, 1.2, 0, 20, 7, 0)        # n_levels = 0
    assert L.nav24_orb_create(C.byref(prm), 0, C.byref(h)) == capi.E_BADARG
    assert L.nav24_orb_sync(None) == capi.E_BADARG
    assert L.nav24_last_error_string(None) == b"null context"


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (never falls back to the oracle)."""
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        pytest.skip("CUDA device present")
    with pytest.raises(capi.Nav24Error) as e:
        capi.OrbContext(1000)
    assert e.value.code == capi.E_CUDA


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nav24_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "orb_oracle" not in txt, f
