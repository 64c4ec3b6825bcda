"""The C-ABI library loads without a GPU and exports every symbol include/hcurl_b200.h declares."""
import ctypes as C
import os
import re

from hephaestus_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "hcurl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(hcurl_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_symbols_are_exported_and_bound(lib_built):
    names = declared_functions()
    assert len(names) > 40
    dll = C.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), "libhcurl_b200.so does not export %s" % n
    missing = [n for n in names if n not in L.SIGNATURES]
    assert not missing, "ctypes binding lacks %s" % missing
    extra = [n for n in L.SIGNATURES if n not in names]
    assert not extra, "bound but not declared in the header: %s" % extra


def test_version_and_error_string_without_gpu(lib_built):
    assert b"sm_100a" in lib_built.hcurl_version()
    assert isinstance(lib_built.hcurl_last_error(), bytes)


def test_no_gpu_means_loud_failure(lib_built):
    """No CPU fallback: without a CUDA device context creation fails with a message."""
    import torch
    if torch.cuda.is_available():
        return
    h = lib_built.hcurl_ctx_create(0, None, 0, 1)
    assert not h
    assert b"CUDA" in lib_built.hcurl_last_error() or b"no CPU fallback" in lib_built.hcurl_last_error()
