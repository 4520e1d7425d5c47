"""Experiment: BLOCK_N choice of the per-tap conv kernel on the encoder's small layers (STATS fprop / PLAIN dgrad)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speak_hack_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
buf = torch.zeros(64 << 20, device=dev)
SHAPES = [(64, 64, 64, 64, 256, 1), (64, 64, 64, 256, 64, 1), (64, 64, 64, 64, 64, 3), (64, 32, 32, 128, 512, 1),
          (64, 32, 32, 512, 128, 1), (64, 32, 32, 128, 128, 3), (64, 16, 16, 256, 1024, 1), (64, 16, 16, 1024, 256, 1),
          (64, 16, 16, 256, 256, 3), (64, 8, 8, 512, 2048, 1), (64, 8, 8, 2048, 512, 1), (64, 8, 8, 512, 512, 3),
          (32, 8, 8, 512, 512, 3), (32, 16, 16, 512, 512, 3)]
for (n, h, w, cin, cout, k) in SHAPES:
    x = torch.randn(n, h, w, cin, device=dev).to(BF)
    wk = (torch.randn(cout, k * k * cin, device=dev) * 0.03).to(BF)
    fl = 2.0 * n * h * w * cin * cout * k * k
    line = f"{cin:4d}->{cout:4d} k{k} @{h:3d} x{n}:"
    for mode, nm in ((ops.EPI_STATS, "stats"), (ops.EPI_PLAIN, "plain")):
        for bn in (64, 128, 256):
            if cout % bn:
                continue
            ts = []
            for _ in range(5):
                buf.add_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv_gemm(x, wk, k, mode, force_block_n=bn)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            line += f"  {nm}{bn}: {sorted(ts)[2] * 1e3:6.1f}us"
    print(line, flush=True)
