"""Experiment: halo-reuse wgrad vs per-tap wgrad (correctness, then speed on the step's 3x3 shapes)."""
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
for (n, h, w, cin, cout) in [(2, 16, 16, 64, 64), (2, 8, 32, 64, 128), (2, 4, 64, 128, 64), (1, 2, 128, 64, 64), (1, 2, 256, 64, 64)]:
    x = torch.randn(n, h, w, cin, generator=g).to(dev).to(BF)
    dy = torch.randn(n, h, w, cout, generator=g).to(dev).to(BF)
    wt = torch.zeros(cout, cin, 3, 3, device=dev, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=1)
    (ref,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    for mode in ("0", "1"):
        os.environ["IRFD_WGRAD_HALO"] = mode
        dw = ops.conv_wgrad(x, dy, 3)
        torch.cuda.synchronize()
        per_tap = [f"{rel(dw[:, :, t // 3, t % 3], ref[:, :, t // 3, t % 3]):.1e}" for t in range(9)]
        print(f"W={w} halo={mode}: rel {rel(dw, ref):.2e}  per tap {per_tap}", flush=True)

buf = torch.zeros(64 << 20, device=dev)
for (n, h, w, cin, cout) in [(32, 256, 256, 128, 64), (32, 256, 256, 64, 64), (32, 128, 128, 256, 128), (32, 128, 128, 128, 128),
                             (64, 64, 64, 64, 64), (64, 32, 32, 128, 128), (64, 16, 16, 256, 256), (32, 64, 64, 256, 256),
                             (32, 32, 32, 512, 512), (32, 16, 16, 512, 512)]:
    x = torch.randn(n, h, w, cin, device=dev).to(BF)
    dy = torch.randn(n, h, w, cout, device=dev).to(BF)
    fl = 2.0 * n * h * w * cin * cout * 9
    for mode in ("0", "1"):
        os.environ["IRFD_WGRAD_HALO"] = mode
        ts = []
        for _ in range(4):
            buf.add_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_wgrad(x, dy, 3)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[1]
        print(f"{cin:4d}->{cout:4d} @{h} x{n}: halo={mode}  {t * 1e3:8.1f} us  {fl / t / 1e9:7.0f} TF/s", flush=True)
    del x, dy
