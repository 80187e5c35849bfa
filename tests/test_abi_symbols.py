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


# ---- the Rust binding cannot be compiled here; keep its extern block honest against the header

C_TO_RUST = {"int": "c_int", "float": "f32", "double": "f64", "uint64_t": "u64", "uint32_t": "u32",
             "int32_t": "i32", "int64_t": "i64", "uint8_t": "u8", "void": "c_void", "char": "c_char"}


def _rust_type_of(c_arg):
    """'const float* x' -> '*const f32'; 'kg_field2d** out' -> '*mut *mut kg_field2d'"""
    c_arg = re.sub(r"/\*.*?\*/", "", c_arg).strip()
    stars = c_arg.count("*")
    words = [w for w in re.sub(r"[*]", " ", c_arg).split()]
    const = "const" in words
    words = [w for w in words if w not in ("const", "struct")]
    base = words[0]
    rust = C_TO_RUST.get(base, base)
    for k in range(stars):
        # only the innermost pointer level carries the C const
        rust = ("*const " if (const and k == 0) else "*mut ") + rust
    return rust


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "krabgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"\b(int|const char\*|uint64_t)\s+(kg_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = [a for a in (x.strip() for x in args.split(",")) if a and a != "void"]
        protos[name] = [_rust_type_of(a) for a in args]
    return protos


def _rust_externs():
    src = open(os.path.join(ROOT, "rust", "src", "engine", "fields", "gpu", "mod.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    block = re.sub(r"//[^\n]*", "", block)
    out = {}
    for name, args in re.findall(r"\bfn\s+(kg_[a-z0-9_]+)\s*\(([^)]*)\)", block):
        out[name] = [a.split(":", 1)[1].strip() for a in args.split(",") if ":" in a]
    return out


def test_rust_extern_block_matches_the_header():
    protos, externs = _header_prototypes(), _rust_externs()
    assert len(externs) >= 30
    unknown = sorted(set(externs) - set(protos))
    assert not unknown, f"bound in rust/ but not declared in krabgpu.h: {unknown}"
    for name, rust_args in externs.items():
        assert rust_args == protos[name], f"{name}: rust {rust_args} != header {protos[name]}"
