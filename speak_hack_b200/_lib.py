"""ctypes binding of libirfd_b200.so (the C ABI declared in include/irfd_b200.h).

There is deliberately no fallback: if the shared library is missing or a launch fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libirfd_b200.so")

IRFD_OK = 0


class IrfdError(RuntimeError):
    pass


# name -> (restype, argtypes); kept in one table so tests can check every symbol of include/irfd_b200.h is exported.
_P = c_void_p
_I = c_int
_F = c_float
_L = c_longlong
SIGNATURES = {
    "irfd_abi_version": (c_int, []),
    "irfd_last_error": (c_char_p, []),
    "irfd_conv_gemm_m_tiles": (c_int, [_I, _I, _I]),
    "irfd_conv_gemm": (c_int, [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "irfd_wgrad_workspace_bytes": (_L, [_I, _I, _I, _I, _I, _I]),
    "irfd_conv_wgrad": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _F, _P, _L, _P]),
}

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
