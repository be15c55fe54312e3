"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from dsopp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(header="dsopp_cuda_pba.h", prefix="dpba_"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    from dsopp_b200 import pose_alignment
    assert header_symbols() == sorted(capi.SIGNATURES)
    assert header_symbols("dsopp_cuda_pose_alignment.h", "dpa_") == sorted(pose_alignment.SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(capi.CUDA_LIB_PATH):
        from dsopp_b200 import build
        build.build_cuda()
    lib = ctypes.CDLL(capi.CUDA_LIB_PATH)
    for name in header_symbols() + header_symbols("dsopp_cuda_pose_alignment.h", "dpa_"):
        assert hasattr(lib, name), name
    assert b"sm_100a" in ctypes.cast(lib.dpba_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DpbaError):
        capi.Handle(2, 16, 64, 64)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "dsopp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert not re.search(r'#include\s+[<"][^>"]*oracle', text), f


def test_every_bound_call_declares_its_argument_types():
    """The bindings pass array addresses as plain integers (capi._ptr): a call without declared argtypes would let ctypes
    truncate them to a C int.  Every dpba_* / dpa_* call must be in its module's SIGNATURES table, every dpbah_* / dpah_*
    call in host.py must have an explicit argtypes line."""
    from dsopp_b200 import pose_alignment
    pkg = os.path.join(ROOT, "dsopp_b200")
    for mod, fname in ((capi, "capi.py"), (pose_alignment, "pose_alignment.py")):
        src = open(os.path.join(pkg, fname)).read()
        called = set(re.findall(r"lib\.((?:dpa|dpba)_[a-z_0-9]+)\(", src))
        assert called <= set(mod.SIGNATURES), sorted(called - set(mod.SIGNATURES))
        assert all(args is not None for _, args in mod.SIGNATURES.values())
    src = open(os.path.join(pkg, "host.py")).read()
    called = set(re.findall(r"lib\.((?:dpbah|dpah)_[a-z_0-9]+)\(", src))
    typed = set(re.findall(r"lib\.((?:dpbah|dpah)_[a-z_0-9]+)\.argtypes", src))
    assert called - typed <= {"dpbah_last_error", "dpah_last_error"}, sorted(called - typed)
