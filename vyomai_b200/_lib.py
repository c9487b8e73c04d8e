"""ctypes binding of libvyom_b200.so, generated from include/vyom_b200.h at import time.

The header is the single source of truth for the C ABI: the struct layouts and function
prototypes below are parsed out of it, so the Python side cannot drift from the library.
There is no CPU fallback — `lib()` raises if the shared library is missing, and every compute
entry point returns VY_ERR_NO_DEVICE without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(REPO_ROOT, "include", "vyom_b200.h")
LIB_PATH = os.path.join(_HERE, "csrc", "libvyom_b200.so")

_CTYPES = {
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "uint32_t": ctypes.c_uint32,
    "uint64_t": ctypes.c_uint64,
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


class VyomError(RuntimeError):
    pass


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return text


def parse_header(path: str = HEADER):
    """Returns (constants, structs, functions) parsed from the C header."""
    src = _strip_comments(open(path).read())
    consts: Dict[str, int] = {}
    for m in re.finditer(r"#define\s+(VY_\w+)\s+(-?\d+)", src):
        consts[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"enum\s*\{([^}]*)\}", src):
        nxt = 0
        for item in m.group(1).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = [s.strip() for s in item.split("=")]
                nxt = int(val, 0)
            else:
                name = item
            consts[name] = nxt
            nxt += 1
    structs: Dict[str, List[Tuple[str, object]]] = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields: List[Tuple[str, object]] = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            is_ptr = "*" in decl
            decl_np = decl.replace("*", " ")
            toks = decl_np.replace(",", " , ").split()
            toks = [t for t in toks if t != "const"]
            base, names = toks[0], [t for t in toks[1:] if t != ","]
            for nm in names:
                fields.append((nm, ctypes.c_void_p if is_ptr else _CTYPES[base]))
        structs[m.group(3)] = fields
    funcs: Dict[str, Tuple[str, List[str]]] = {}
    for m in re.finditer(r"VY_API\s+([\w\s\*]+?)\s*\b(vy_\w+)\s*\(([^)]*)\)\s*;", src):
        ret = " ".join(m.group(1).split())
        args = [a.strip() for a in m.group(3).split(",") if a.strip() and a.strip() != "void"]
        funcs[m.group(2)] = (ret, args)
    return consts, structs, funcs


CONSTS, _STRUCT_FIELDS, FUNCS = parse_header()
globals().update(CONSTS)

STRUCTS: Dict[str, type] = {}
for _name, _fields in _STRUCT_FIELDS.items():
    STRUCTS[_name] = type(_name, (ctypes.Structure,), {"_fields_": _fields})

_lib = None


def _restype(ret: str):
    if ret == "const char*" or ret == "const char *":
        return ctypes.c_char_p
    if "*" in ret:
        return ctypes.c_void_p
    return _CTYPES[ret]


def lib() -> ctypes.CDLL:
    """Loads the C-ABI library once. Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VyomError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C vyomai_b200/csrc`). vyomai_b200 has no CPU or PyTorch fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    for name, (ret, args) in FUNCS.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = _restype(ret)
        fn.argtypes = [(ctypes.c_char_p if "char" in a else ctypes.c_void_p) if "*" in a else _CTYPES[a.split()[0]]
                       for a in args]
    if L.vy_version() != CONSTS["VY_ABI_VERSION"]:
        raise VyomError("libvyom_b200.so ABI version does not match include/vyom_b200.h")
    for sname, cls in STRUCTS.items():
        have = L.vy_abi_sizeof(sname.encode())
        if have != ctypes.sizeof(cls):
            raise VyomError(
                f"libvyom_b200.so is stale: sizeof({sname}) is {have} in the library but {ctypes.sizeof(cls)} in "
                "include/vyom_b200.h — rebuild with `make -C vyomai_b200/csrc`"
            )
    _lib = L
    return L


def last_error() -> str:
    return lib().vy_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise VyomError(f"{what} failed ({rc}): {last_error()}")


class KernelTimer:
    """Optional per-call CUDA-event timing of the C-ABI entry points (bench.py's roofline leg).
    Events are recorded on the stream the kernel is launched on; nothing synchronises until
    `summary()`."""

    def __init__(self):
        self.records = []  # (fn_name, info dict, start event, end event)

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, info, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["calls"] += 1
            d["ms"] += ms
            d["flops"] += info.get("flops", 0.0)
            d["bytes"] += info.get("bytes", 0.0)
        return out


TIMER: "KernelTimer | None" = None


def _work(fn_name: str, kw: dict) -> dict:
    """Algorithmic work of one call: GEMM FLOPs = 2 M N K; attention FLOPs =
    4 B Hq Sq Skv 64 per matmul pair (halved when causal); the rest are byte counts."""
    try:
        if fn_name == "vy_gemm":
            return {"flops": 2.0 * kw["M"] * kw["N"] * kw["K"]}
        if fn_name in ("vy_attn_fwd", "vy_attn_bwd"):
            f = 4.0 * kw["B"] * kw["n_q_heads"] * kw["Sq"] * kw["Skv"] * 64
            if kw.get("causal"):
                f *= 0.5
            return {"flops": f * (3.5 if fn_name == "vy_attn_bwd" else 1.0)}
        if fn_name == "vy_attn_decode":
            es = 2 if kw["cache_dtype"] == CONSTS["VY_BF16"] else 4
            return {"bytes": 2.0 * kw["B"] * kw["n_kv_heads"] * (kw["start_pos"] + 1) * 64 * es}
        if fn_name in ("vy_add_layernorm_fwd", "vy_add_layernorm_bwd"):
            es = 2 if kw["io_dtype"] == CONSTS["VY_BF16"] else 4
            n = 2 + (1 if kw.get("residual") else 0) + (1 if kw.get("sum_out") else 0)
            if fn_name.endswith("bwd"):
                n = 3
            return {"bytes": float(n) * kw["rows"] * kw["H"] * es}
    except KeyError:
        pass
    return {}


def _timed(fn_name: str, kw: dict, thunk):
    t = TIMER
    if t is None:
        return thunk()
    import torch
    stream = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    rc = thunk()
    e1.record(stream)
    t.records.append((fn_name, _work(fn_name, kw), e0, e1))
    return rc


def call(fn_name: str, struct_name: str, **kw) -> None:
    """Fills `struct_name` from keyword arguments and calls `fn_name(&struct)`."""
    st = STRUCTS[struct_name]()
    for k, v in kw.items():
        if v is None:
            continue
        setattr(st, k, v)
    fn = getattr(lib(), fn_name)
    rc = _timed(fn_name, kw, lambda: fn(ctypes.byref(st)))
    check(rc, fn_name)


def call_raw(fn_name: str, *args) -> None:
    """Calls a scalar-argument entry point (vy_colsum, vy_argmax_rows, ...)."""
    fn = getattr(lib(), fn_name)
    rc = _timed(fn_name, {}, lambda: fn(*args))
    check(rc, fn_name)
