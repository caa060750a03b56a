"""The C-ABI library loads on a CPU-only box and exports every symbol include/pfem_b200.h declares."""
import ctypes
import os
import re

from pfem_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "pfem_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pfem_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.SYMBOLS), "include/pfem_b200.h and pfem_b200/capi.py list different entry points"


def test_library_exports_every_symbol():
    assert os.path.exists(capi.LIB_PATH), "libpfem_b200.so missing: run __graft_entry__.build()"
    L = ctypes.CDLL(capi.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    L.pfem_abi_version.restype = ctypes.c_int
    assert L.pfem_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """Without a CUDA device pfem_create must fail loudly (there is no CPU path)."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        capi.PfemContext(3, 0)
    except capi.PfemError as e:
        assert e.code == -2 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("pfem_create succeeded without a GPU")
