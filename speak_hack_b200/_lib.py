"""ctypes binding of libirfd_b200.so (the C ABI declared in include/irfd_b200.h).

The signature table is parsed from the header itself, so the Python binding cannot drift from the declared ABI and
tests can assert that every declared symbol is exported.  There is deliberately no fallback: if the shared library is
missing or a launch fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libirfd_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "irfd_b200.h")

IRFD_OK = 0


class IrfdError(RuntimeError):
    pass


_CTYPES = {
    "int": c_int,
    "float": c_float,
    "long long": c_longlong,
    "irfd_stream_t": c_void_p,
    "const char*": c_char_p,
}


def _ctype(decl: str):
    t = decl.strip()
    if "*" in t:
        return c_char_p if t.replace(" ", "") == "constchar*" else c_void_p
    return _CTYPES[t]


def parse_header(path: str = HEADER_PATH):
    """Return {name: (restype, [argtypes])} for every function prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    src = src.replace('extern "C" {', "")
    src = re.sub(r"typedef[^;]*;", "", src)
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(irfd_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                # strip the parameter name (last identifier) keeping the type
                tm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_ctype(tm.group(1)))
        sigs[name] = (_ctype(ret), argtypes)
    return sigs


SIGNATURES = parse_header()

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises IrfdError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IrfdError(
            f"{LIB_PATH} not found: build it with `python -m speak_hack_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU/PyTorch fallback for the IRFD hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != IRFD_OK:
        msg = load().irfd_last_error()
        raise IrfdError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
