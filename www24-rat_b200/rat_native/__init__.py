"""ctypes binding of librat_b200.so (the C ABI declared in include/rat_b200.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the ABI.  There is no
fallback: if the library is missing or the current device is not sm_100 every call raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
HEADER = os.path.join(os.path.dirname(_PKG), "include", "rat_b200.h")
LIB_PATH = os.path.join(_PKG, "lib", "librat_b200.so")

_SCALARS = {
    "int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float, "double": ctypes.c_double,
    "unsigned long long": ctypes.c_ulonglong, "unsigned int": ctypes.c_uint, "size_t": ctypes.c_size_t,
}
_RET = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "const char*": ctypes.c_char_p,
        "long long": ctypes.c_longlong}


class RatError(RuntimeError):
    pass


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """{name: (return_type, [(ctype, argname), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"(const char\*|int|size_t|long long)\s+(rat_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                typ = mm.group(1).strip()
                parsed.append((typ, mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


def _ctype(typ: str):
    if typ.endswith("*"):
        return ctypes.c_void_p
    return _SCALARS[typ]


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RatError(f"{LIB_PATH} not found: build it with `python www24-rat_b200/build.py` "
                           f"(there is no CPU fallback for the RAT hot path)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.fn = {}
        for name, (ret, args) in self.protos.items():
            f = getattr(self.cdll, name)          # AttributeError => header/library drift
            f.restype = _RET[ret]
            f.argtypes = [_ctype(t) for t, _ in args]
            self.fn[name] = f
        if self.fn["rat_abi_version"]() != 1:
            raise RatError("librat_b200.so ABI version mismatch")

    def last_error(self) -> str:
        return self.fn["rat_last_error"]().decode()


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def _ptr(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x


_profile = None      # {name: [(start_event, end_event), ...]} while profiling


def profile_calls(enable: bool):
    """Record a CUDA-event pair around every entry-point call (bench.py's per-kernel timing pass)."""
    global _profile
    _profile = {} if enable else None


def profile_results():
    """{name: (n_calls, total_ms)}; call after torch.cuda.synchronize()."""
    out = {}
    for name, pairs in (_profile or {}).items():
        out[name] = (len(pairs), sum(a.elapsed_time(b) for a, b in pairs))
    return out


def call(name: str, *args):
    """Call an int-returning entry point with torch tensors / scalars; raise RatError on failure."""
    L = lib()
    f = L.fn[name]
    if _profile is not None:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = f(*[_ptr(x) for x in args])
        b.record()
        _profile.setdefault(name, []).append((a, b))
    else:
        rc = f(*[_ptr(a) for a in args])
    if L.protos[name][0] == "int" and rc != 0 and not name.endswith(("_blocks", "_version")):
        raise RatError(f"{name} failed (rc={rc}): {L.last_error()}")
    return rc


def query(name: str, *args):
    """Call a size/int query (no error convention)."""
    return lib().fn[name](*[_ptr(a) for a in args])


def require_device():
    rc = lib().fn["rat_device_check"]()
    if rc != 0:
        raise RatError("RAT hot path needs an sm_100 (B200) CUDA device: " + lib().last_error())


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
