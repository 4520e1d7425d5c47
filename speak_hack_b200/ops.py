"""Tensor-level wrappers over the C ABI: validate, allocate outputs with torch (device memory plumbing), launch.

Nothing in here computes with PyTorch; every function either launches a kernel from libirfd_b200.so or raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32

EPI_PLAIN, EPI_STATS, EPI_STYLE = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.IrfdError(f"{name}: expected a CUDA tensor (the IRFD hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.IrfdError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.IrfdError(f"{name}: expected a contiguous tensor")


def conv_gemm(
    x: torch.Tensor,            # [N,H,W,Cin] bf16
    wk: torch.Tensor,           # [Cout, k*k*Cin] bf16 (tap-major)
    ksize: int,
    mode: int = EPI_PLAIN,
    bias: Optional[torch.Tensor] = None,
    nw: Optional[torch.Tensor] = None,
    noise: Optional[torch.Tensor] = None,
    sp1: Optional[torch.Tensor] = None,
    s1: Optional[torch.Tensor] = None,
    force_block_n: int = 0,
):
    """Stride-1 same-padding conv as implicit GEMM.  Returns out (mode 0), (out, sum, sq) (mode 1), (a, y) (mode 2)."""
    lib = _lib.load()
    _chk(x, BF16, "x")
    _chk(wk, BF16, "wk")
    n, h, w, cin = x.shape
    cout = wk.shape[0]
    if wk.shape[1] != ksize * ksize * cin:
        raise _lib.IrfdError(f"wk shape {tuple(wk.shape)} does not match ksize={ksize}, cin={cin}")
    out = torch.empty((n, h, w, cout), dtype=BF16, device=x.device)
    out2 = torch.empty_like(out) if mode == EPI_STYLE else None
    ssum = ssq = None
    if mode == EPI_STATS:
        mt = lib.irfd_conv_gemm_m_tiles(n, h, w)
        ssum = torch.empty((mt, cout), dtype=F32, device=x.device)
        ssq = torch.empty((mt, cout), dtype=F32, device=x.device)
    for t, nm in ((bias, "bias"), (nw, "nw"), (noise, "noise"), (sp1, "sp1"), (s1, "s1")):
        if t is not None:
            _chk(t, F32, nm)
    rc = lib.irfd_conv_gemm(
        x.data_ptr(), n, h, w, cin, wk.data_ptr(), cout, ksize, out.data_ptr(), _ptr(out2), mode, _ptr(bias), _ptr(nw),
        _ptr(noise), _ptr(sp1), _ptr(s1), _ptr(ssum), _ptr(ssq), force_block_n, _stream(),
    )
    _lib.check(rc, "irfd_conv_gemm")
    if mode == EPI_STATS:
        return out, ssum, ssq
    if mode == EPI_STYLE:
        return out, out2
    return out


_workspace = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only per-device scratch buffer (split-K partials etc.); stream-ordered use only."""
    key = (device.type, device.index)
    buf = _workspace.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspace[key] = buf
    return buf


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, ksize: int, dw: Optional[torch.Tensor] = None, beta: float = 0.0):
    """dW (OIHW fp32) of a stride-1 same-padding conv from its NHWC bf16 input and output gradient."""
    lib = _lib.load()
    _chk(x, BF16, "x")
    _chk(dy, BF16, "dy")
    n, h, w, cin = x.shape
    cout = dy.shape[-1]
    if dw is None:
        dw = torch.empty((cout, cin, ksize, ksize), dtype=F32, device=x.device)
        beta = 0.0
    _chk(dw, F32, "dw")
    need = lib.irfd_wgrad_workspace_bytes(n, h, w, cin, cout, ksize)
    ws = workspace(need, x.device)
    rc = lib.irfd_conv_wgrad(x.data_ptr(), dy.data_ptr(), n, h, w, cin, cout, ksize, dw.data_ptr(), beta,
                             ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "irfd_conv_wgrad")
    return dw
