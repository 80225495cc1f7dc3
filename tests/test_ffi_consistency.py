"""Bindings above the C ABI, checked textually on CPU.  The Rust crate (rust/) cannot be compiled in this image, so
every `extern "C"` declaration of rust/src/ffi.rs must name an entry point of include/ndi_b200.h with the same
number of parameters and ABI-compatible parameter types, and every NDI_* constant must carry the header's value;
the ctypes table the test-suite calls through is held to the same check."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    text = open(os.path.join(ROOT, "include", "ndi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    code = re.sub(r"^\s*#[^\n]*$", " ", text, flags=re.M)          # prototypes only: no preprocessor lines
    for ret, name, args in re.findall(r"([\w\s\*]+?)\b(ndi_\w+)\s*\(([^;{}]*?)\)\s*;", code):
        args = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []
        protos[name] = (ret.strip(), args)
    consts = {k: v for k, v in re.findall(r"#define\s+(NDI_\w+)\s+(\d+)u?\b", text)}
    return protos, consts


def _c_kind(arg):
    """ABI class of a C parameter declaration"""
    t = re.sub(r"\b\w+$", "", arg).strip() if not arg.endswith("*") else arg      # drop the parameter name
    t = t.replace("const ", "").strip()
    if "*" in t:
        return "ptr"
    return {"int32_t": "i32", "ndi_status": "i32", "ndi_dtype": "i32", "int64_t": "i64", "uint32_t": "u32",
            "uint64_t": "u64"}[t]


def _rust_kind(t):
    t = t.strip()
    if t.startswith("*"):
        return "ptr"
    return {"i32": "i32", "ndi_status": "i32", "i64": "i64", "u32": "u32", "u64": "u64"}[t]


def _rust():
    text = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    fns = {}
    for name, args, ret in re.findall(r"pub fn (ndi_\w+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", text):
        params = [a.split(":", 1)[1] for a in args.split(",") if ":" in a]
        fns[name] = (params, (ret or "").strip())
    consts = {k: v for k, v in re.findall(r"pub const (NDI_\w+)\s*:\s*\w+\s*=\s*(\d+)\s*;", text)}
    return fns, consts


def test_every_rust_declaration_matches_the_header():
    protos, _ = _header()
    fns, _ = _rust()
    assert len(fns) >= 15
    for name, (params, ret) in fns.items():
        assert name in protos, f"{name} is declared in ffi.rs but not in include/ndi_b200.h"
        c_ret, c_args = protos[name]
        assert len(params) == len(c_args), (name, params, c_args)
        for p, a in zip(params, c_args):
            assert _rust_kind(p) == _c_kind(a), (name, p, a)
        if "char" in c_ret:
            assert ret.startswith("*const c_char"), (name, ret)
        else:
            assert ret == "ndi_status", (name, ret)


def test_rust_constants_carry_the_header_values():
    _, c = _header()
    _, r = _rust()
    assert {"NDI_OK", "NDI_OUT_OF_BOUNDS", "NDI_NAN_QUERY", "NDI_PERIODIC_MISMATCH", "NDI_NOT_MONOTONIC", "NDI_F32",
            "NDI_F64", "NDI_I32", "NDI_I64", "NDI_U32", "NDI_U64", "NDI_ASSUME_VALID"} <= set(r)
    for k, v in r.items():
        assert c.get(k) == v, (k, v, c.get(k))


def test_rust_build_script_compiles_every_cuda_source():
    text = open(os.path.join(ROOT, "rust", "build.rs")).read()
    from ndarray_interp_b200.build import SOURCES
    for s in SOURCES:
        assert f'"{s}"' in text, s
    assert "arch=compute_100a,code=sm_100a" in text and "-fmad=false" in text


def test_python_ctypes_table_matches_the_header_parameter_by_parameter():
    """the table the test-suite itself calls through: same arity and the same ABI class per parameter"""
    import ctypes as C

    from ndarray_interp_b200 import _lib
    protos, _ = _header()

    def ct_kind(t):
        if t in (C.c_int32,):
            return "i32"
        if t in (C.c_int64,):
            return "i64"
        if t in (C.c_uint32,):
            return "u32"
        if t in (C.c_uint64,):
            return "u64"
        return "ptr"                       # c_void_p, POINTER(...)

    assert set(_lib.SIGNATURES) == set(protos)
    for name, (res, args) in _lib.SIGNATURES.items():
        c_ret, c_args = protos[name]
        assert len(args) == len(c_args), (name, len(args), c_args)
        for t, a in zip(args, c_args):
            assert ct_kind(t) == _c_kind(a), (name, t, a)
        assert (res is C.c_char_p) == ("char" in c_ret) and (res is C.c_int32) == (c_ret in ("ndi_status",)) or \
            (res is C.c_uint64 and c_ret == "uint64_t"), (name, res, c_ret)
