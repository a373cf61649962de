"""CPU-side checks of the C-ABI: the library builds/loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "frozenbilm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from frozenbilm_b200.build import build_library
    from frozenbilm_b200 import lib
    path = build_library()
    dll = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in include/frozenbilm_b200.h but not exported"
    assert sorted(lib.EXPORTS) == declared
    dll.fb_abi_version.restype = ctypes.c_int
    assert dll.fb_abi_version() == 3
    # the hardware probe / phase-stamp hooks live only in the test-only debug library
    for name in lib.DEBUG_EXPORTS:
        assert not hasattr(dll, name), f"{name} must not be exported by the product library"
    dbg = ctypes.CDLL(lib.LIB_DBG_PATH)
    dbg_hdr = open(os.path.join(ROOT, "include", "frozenbilm_b200_debug.h")).read()
    dbg_hdr = re.sub(r"/\*.*?\*/", "", dbg_hdr, flags=re.S)
    assert sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", dbg_hdr))) == sorted(lib.DEBUG_EXPORTS)
    for name in declared + lib.DEBUG_EXPORTS:
        assert hasattr(dbg, name)


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from frozenbilm_b200 import lib
    l = lib.load()
    args = lib.GemmArgs()
    args.M = args.N = args.K = 128
    args.batch = 1
    rc = l.fb_gemm(ctypes.byref(args), None)
    assert rc != 0 and l.fb_last_error()
