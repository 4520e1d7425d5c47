"""GPU micro-benchmark: what bounds the small-K (HBM-bound) 1x1 encoder GEMMs?  STATS vs PLAIN epilogue, BLOCK_N 64/128/256."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speak_hack_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
buf = torch.zeros(64 << 20, dtype=torch.float32, device=dev)


def timeit(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        buf.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


g = torch.Generator().manual_seed(0)
for name, h, cin, cout in (("64->256 @64", 64, 64, 256), ("256->64 @64", 64, 256, 64), ("128->512 @32", 32, 128, 512),
                           ("256->1024 @16", 16, 256, 1024), ("1024->256 @16", 16, 1024, 256)):
    x = (torch.randn(192, h, h, cin, generator=g) * 0.5).to(dev).to(torch.bfloat16)
    wk = (torch.randn(3 * cout, cin, generator=g) * 0.05).to(dev).to(torch.bfloat16)
    m = 192 * h * h
    nbytes = 2.0 * m * (cin + cout)
    line = [f"{name:16s}"]
    for mode, mn in ((ops.EPI_PLAIN, "plain"), (ops.EPI_STATS, "stats")):
        for bn in (0, 64, 128, 256):
            if bn and cout % bn:
                continue
            t = timeit(lambda: ops.conv_gemm_grouped(x, wk, 1, mode, wgroups=3, force_block_n=bn))
            line.append(f"{mn}/bn{bn}: {t:6.1f} us ({nbytes / t / 1e3:5.0f} GB/s)")
    print("  ".join(line))
