"""Experiment: halo-reuse conv kernel vs per-tap kernel (correctness of the shifted-descriptor trick, then speed)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speak_hack_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


g = torch.Generator().manual_seed(0)
n, h, w, cin, cout = 2, 8, 128, 128, 64
x = torch.randn(n, h, w, cin, generator=g).to(dev).to(BF)
wt = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev).to(BF)
wk = wt.permute(0, 2, 3, 1).contiguous().reshape(cout, 9 * cin)
ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
for mode in ("0", "1", "2"):
    os.environ["IRFD_CONV_HALO"] = mode
    try:
        y = ops.conv_gemm(x, wk, 3, ops.EPI_PLAIN)
        torch.cuda.synchronize()
        print(f"IRFD_CONV_HALO={mode}: rel_l2 vs torch = {rel(y.float(), ref):.3e}", flush=True)
    except Exception as exc:  # noqa: BLE001
        print(f"IRFD_CONV_HALO={mode}: FAILED {exc}", flush=True)
        sys.exit(1)

buf = torch.zeros(64 << 20, device=dev)
for (n, h, w, cin, cout) in [(32, 256, 256, 128, 64), (32, 256, 256, 64, 64), (32, 256, 256, 64, 128),
                             (32, 128, 128, 256, 128), (32, 128, 128, 128, 128)]:
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(BF) if False else torch.randn(n, h, w, cin, device=dev).to(BF)
    wk = (torch.randn(cout, 9 * cin, device=dev) * 0.03).to(BF)
    fl = 2.0 * n * h * w * cin * cout * 9
    for mode in ("0", "1"):
        os.environ["IRFD_CONV_HALO"] = mode
        ts = []
        for _ in range(4):
            buf.add_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_gemm(x, wk, 3, ops.EPI_PLAIN)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[1]
        print(f"{cin:4d}->{cout:4d} @{h}: halo={mode}  {t * 1e3:8.1f} us  {fl / t / 1e9:7.0f} TF/s", flush=True)
    del x
