"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/tribe_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

import algonauts2025_b200
from algonauts2025_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tribe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tribe_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == _lib.EXPORTS


def test_library_builds_loads_and_exports_every_symbol():
    path = algonauts2025_b200.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/tribe_b200.h but not exported"
    lib.tribe_abi_version.restype = ctypes.c_int
    assert lib.tribe_abi_version() == 1


def test_gemm_struct_layout_matches_header():
    """ctypes mirrors of the ABI structs: field order and sizes must match the C declaration."""
    text = open(os.path.join(ROOT, "include", "tribe_b200.h")).read()
    body = re.search(r"typedef struct TribeGemm \{(.*?)\} TribeGemm;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
    assert names == [f[0] for f in _lib.TribeGemm._fields_]
    assert ctypes.sizeof(_lib.TribeOperand) == 72


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: the product path fails loudly without CUDA tensors."""
    import torch

    from algonauts2025_b200 import ops

    with pytest.raises(algonauts2025_b200.TribeError):
        ops.scalenorm_fwd(torch.zeros(2, 8), torch.ones(1), torch.zeros(2, 8, dtype=torch.bfloat16), torch.zeros(2))
