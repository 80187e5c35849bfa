"""CPU-side checks of the drop-in boundary: libkrabgpu.so loads, exports every symbol that
include/krabgpu.h declares, and refuses to run without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import krabmaga_b200 as kb
from krabmaga_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "krabgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kg_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    so = abi.build()
    L = C.CDLL(so)
    syms = header_symbols()
    assert len(syms) >= 45
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in krabgpu.h but not exported: {missing}"


def test_python_binding_covers_the_header():
    L = abi.lib()
    assert sorted(L._declared) == header_symbols()
    assert L.kg_abi_version() == 1


def test_params_struct_layout_matches_header():
    # 7 floats + i32 + 2 x u64 = 48 bytes, natural alignment
    assert C.sizeof(abi.KgBoidsParams) == 48
    assert abi.KgBoidsParams.seed.offset == 32 and abi.KgBoidsParams.step.offset == 40


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(kb.KgError) as e:
        kb.Field2D(10.0, 10.0, 0.5, True)
    assert e.value.code == abi.KG_E_CUDA
    with pytest.raises(kb.KgError):
        kb.DenseNumberGrid2D(8, 8)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under krabmaga_b200/ may reference it."""
    pkg = os.path.join(ROOT, "krabmaga_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"oracle_binding|liboracle|oracle/|krabmaga_oracle", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
